#!/usr/bin/env python
"""
Fixture for BASELINE config 2's parity clause ("per-window probs vs oracle on a fixed 2,048-window subsample", batch 1024):

  * 2,048 windows of the counter-based stream (genomad_b200/synth.py, seed 1; incl. >= 32 windows of the 1 % N-run / IUPAC
    sub-stream) through the CPU oracle (oracle/igloo_model.py::forward, fp32) with the SHIPPED weights;
  * 320 windows (256 of the same stream + 64 from the 8 worst-case families of tools/precision_study.py) with SYNTHETIC
    O(1) IGLOO weights (oracle.igloo_model.synthetic_igloo_weights) -- live patch gather, logits GEMM and softmax.

Only indices / seeds and the oracle's outputs are stored (the windows are regenerated from the counter):

    python tests/golden/make_config2_golden.py      # ~3 min on 8 cores -> tests/golden/config2_subsample.npz
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from genomad_b200 import synth  # noqa: E402
from oracle import igloo_model as M, tokenizer as T  # noqa: E402
import precision_study  # noqa: E402


def oracle(a, w, dtype=torch.float32, bs=32):
    tok = T.tokenize_windows(a)
    return np.concatenate([M.forward(tok[i:i + bs], w, dtype) for i in range(0, len(tok), bs)])


def main():
    w = M.load_npz_weights(ROOT / "genomad_b200" / "data" / "nn_classifier.npz")
    wsyn = M.synthetic_igloo_weights(w)
    idx = synth.subsample_indices(2048, 1_000_000, seed=1)
    a = synth.windows_numpy(idx, seed=1)
    n_dirty = int((a == ord("N")).any(1).sum())
    assert n_dirty >= 32, n_dirty
    shipped32 = oracle(a, w)
    idx_syn = idx[::8][:256]
    fam = precision_study.make_windows(64, seed=101)
    a_syn = np.concatenate([synth.windows_numpy(idx_syn, seed=1), fam])
    syn32 = oracle(a_syn, wsyn)
    syn64 = oracle(a_syn, wsyn, torch.float64)
    np.savez_compressed(ROOT / "tests" / "golden" / "config2_subsample.npz", indices=idx, shipped_fp32=shipped32,
                        syn_indices=idx_syn, syn_family_seed=np.int64(101), synthetic_fp32=syn32, synthetic_fp64=syn64,
                        n_dirty=np.int64(n_dirty))
    print("windows", len(idx), "dirty", n_dirty, "| synthetic", len(a_syn),
          "| fp32 vs fp64 (synthetic):", float(np.abs(syn32 - syn64).max()))


if __name__ == "__main__":
    main()
