#!/usr/bin/env python
"""
BASELINE config 5's batch axis as a markdown table: runs `bench.py --config 5` (batch 128 ... 4096 per GPU, windows/s and
per-kernel roofline fractions from the library's per-stage CUDA events) and formats its `sweep` rows.

    python tools/batch_sweep.py > gpurun_out/r02_batch_sweep.md
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--config", "5"], capture_output=True, text=True, check=True)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    print("| batch | ms/step | windows/s | conv2 frac of sustained bf16 (algorithmic, ceiling 0.5) | w_v + gather frac of HBM | layer-1 frac of HBM | small kernels ms | stage ms |")
    print("|---|---|---|---|---|---|---|---|")
    for row in line["sweep"]:
        st = "  ".join(f"{k}={v:.3f}" for k, v in row["stage_ms"].items())
        wg = row["wv_gather_frac_of_hbm"]
        print(f"| {row['batch']} | {row['ms_per_step']:.3f} | {row['windows_per_s']:,.0f} | {row['conv2_frac_of_bf16_sustained']:.2f} | "
              f"{wg:.2f} | {row['layer1_frac_of_hbm']:.2f} | {row['small_kernels_ms']:.3f} | {st} |" if wg is not None else
              f"| {row['batch']} | {row['ms_per_step']:.3f} | {row['windows_per_s']:,.0f} | {row['conv2_frac_of_bf16_sustained']:.2f} | - | "
              f"{row['layer1_frac_of_hbm']:.2f} | {row['small_kernels_ms']:.3f} | {st} |")


if __name__ == "__main__":
    main()
