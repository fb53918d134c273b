#!/usr/bin/env python
"""
BASELINE config 5's batch sweep: windows/s and per-kernel roofline fractions at batch 256 ... 4096 on one B200.
Prints a markdown table (copied to profiles/).  Peaks from MEASURED_PEAKS.json (sustained bf16 TF/s, HBM GB/s).
    python tools/batch_sweep.py [batch ...]
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from genomad_b200.engine import Classifier

FLOP_CONV = 2 * 5997 * 6 * 128 * 128            # one causal Conv1D 128->128 k=6, per window
FLOP_WV = 2 * 5997 * 128 * 128


def main():
    batches = [int(x) for x in sys.argv[1:]] or [256, 512, 1024, 2048, 4096]
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    tf_peak = float(pk.get("bf16_tflops_sustained", pk.get("bf16_tflops", 1455.4)))
    hbm_peak = float(pk.get("hbm_gbps", 6572.5))
    print(f"peaks: {tf_peak:.0f} TF/s (bf16 sustained), {hbm_peak:.0f} GB/s HBM\n")
    print("| batch | ms/step | windows/s | conv2 ms (alg. TF/s, frac) | conv3 ms | w_v ms (GB/s, frac) | gather ms (GB/s, frac) | layer-1 ms (GB/s, frac) | small kernels ms |")
    print("|---|---|---|---|---|---|---|---|---|")
    g = torch.Generator(device="cuda").manual_seed(1)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    for B in batches:
        clf = Classifier(max_batch=B)
        pool = [acgt[torch.randint(0, 4, (B, 6000), device="cuda", generator=g)] for _ in range(3)]
        out = torch.empty((B, 3), device="cuda")
        for i in range(3):
            clf.predict_ascii(pool[i % 3], out)
        torch.cuda.synchronize()
        K = max(8, 20480 // B)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            clf.predict_ascii(pool[i % 3], out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / K
        clf.set_option("profile_stages", 1)
        reps = 4
        for i in range(reps):
            clf.predict_ascii(pool[i % 3], out)
        torch.cuda.synchronize()
        acc = {}
        for name, t in clf.stage_times():
            acc[name] = acc.get(name, 0.0) + t / reps
        clf.set_option("profile_stages", 0)
        conv2, conv3 = acc["conv2"], acc["conv3"]
        wv = (acc["wv0"] + acc["wv1"]) / 2
        ga = (acc["gather0"] + acc["gather1"]) / 2
        emb = acc["embed_conv1"]
        small = sum(v for k, v in acc.items() if k.startswith(("logits", "attention", "head")))
        tf = B * FLOP_CONV / (conv2 * 1e-3) / 1e12
        wv_gbs = B * (5997 * 512 + 749 * 512) / (wv * 1e-3) / 1e9
        ga_gbs = B * (4510 * 512 + 8880 * 4) / (ga * 1e-3) / 1e9          # distinct rows only (duplicates are L1 hits)
        emb_gbs = B * (6000 + 5997 * 768) / (emb * 1e-3) / 1e9
        print(f"| {B} | {ms:.2f} | {B / ms * 1e3:,.0f} | {conv2:.2f} ({tf:.0f}, {tf / tf_peak:.2f}) | {conv3:.2f} | "
              f"{wv:.2f} ({wv_gbs:.0f}, {wv_gbs / hbm_peak:.2f}) | {ga:.2f} ({ga_gbs:.0f}, {ga_gbs / hbm_peak:.2f}) | "
              f"{emb:.2f} ({emb_gbs:.0f}, {emb_gbs / hbm_peak:.2f}) | {small:.2f} |", flush=True)
        del clf, pool, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
