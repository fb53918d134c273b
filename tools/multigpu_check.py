#!/usr/bin/env python
"""
Multi-GPU check, run under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py

Long contigs (BASELINE config 4 shape, shortened) are sharded across ranks so contigs straddle the shard boundaries;
the distributed result must equal the single-GPU result bitwise in "gather" mode and to 1e-6 in "allreduce" mode.
Also runs the module driver end to end (rank 0 writes the outputs).
"""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from genomad_b200 import dist as gdist, engine, nn_classification, sequence  # noqa: E402


def main():
    info = gdist.init_process_group_if_needed("nccl")
    assert info.world_size > 1, "run under torchrun with >= 2 processes"
    rng = np.random.default_rng(0)                         # same FASTA on every rank
    tmp = Path(tempfile.gettempdir()) / "gnm_multigpu"
    if info.is_main:
        tmp.mkdir(exist_ok=True)
        with open(tmp / "long.fna", "w") as fh:
            for i, ln in enumerate([200_000, 61_000, 3_000, 149_999, 6_000, 300_500, 2_499]):
                s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, ln)].tobytes().decode()
                fh.write(f">long_{i} len={ln}\n{s}\n")
    dist.barrier()
    fa = tmp / "long.fna"
    enc = sequence.encode_fasta(fa)
    n = enc.windows.shape[0]
    clf = engine.Classifier(None, device=info.local_rank, max_batch=64)
    dev = torch.device("cuda", info.local_rank)
    # single-GPU reference, computed redundantly on every rank
    full = torch.from_numpy(clf.classify_host(enc.windows)).to(dev)
    ref = clf.segment_mean(full, torch.from_numpy(enc.offsets).to(dev)).cpu().numpy()
    got_g = nn_classification._classify_windows(clf, enc.windows, enc.offsets, info, "gather")
    got_a = nn_classification._classify_windows(clf, enc.windows, enc.offsets, info, "allreduce")
    s, e = gdist.shard_bounds(n, info.world_size, info.rank)
    straddle = int(((enc.offsets[:-1] < e) & (enc.offsets[1:] > s) & ((enc.offsets[:-1] < s) | (enc.offsets[1:] > e))).sum())
    ok_g = np.array_equal(got_g, ref)
    err_a = float(np.abs(got_a - ref).max())
    print(f"[rank {info.rank}] windows {n} shard [{s},{e}) contigs straddling my shard: {straddle}  "
          f"gather bitwise-equal: {ok_g}  allreduce max|d|: {err_a:.2e}", flush=True)
    assert ok_g and err_a < 1e-6
    # module driver end to end, both contig reducers (gather: bitwise; allreduce of partial sums: fp32 re-association only)
    out = tmp / "out"
    for reducer, tol in (("gather", 0.0), ("allreduce", 1e-6)):
        nn_classification.main(fa, out, False, 64, True, 4, False, True, contig_reduce=reducer)
        dist.barrier()
        if info.is_main:
            z = np.load(out / "long_nn_classification" / "long_nn_classification.npz")
            err = float(np.abs(z["predictions"] - ref).max())
            assert err <= tol, f"module output ({reducer}) differs from the single-GPU result by {err}"
            what = "NPZ equals single-GPU result bitwise" if tol == 0.0 else f"max |d| vs single GPU {err:.2e}"
            print(f"module driver under torchrun x{info.world_size} ({reducer}): {what};",
                  (out / "long_nn_classification" / "long_nn_classification.tsv").read_text().splitlines()[1], flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
