#!/usr/bin/env python
"""
FASTA -> TSV wall clock of the module driver on one GPU (SURVEY.md section 8(d) "module" metric).

    python tools/e2e_fasta_bench.py [n_contigs] [batch_size]

Writes a synthetic metagenome-like FASTA (contig lengths log-uniform in [1 kb, 120 kb], 60-column lines), runs
genomad_b200.nn_classification.main() twice (second run with --restart so the file cache is warm) and prints the
phase times parsed from the log plus windows/s over the whole call.
"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from genomad_b200 import nn_classification, sequence  # noqa: E402


def main():
    n_contigs = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    rng = np.random.default_rng(0)
    tmp = Path(tempfile.mkdtemp(prefix="gnm_e2e_"))
    fa = tmp / "meta.fna"
    lens = np.exp(rng.uniform(np.log(1000), np.log(120000), n_contigs)).astype(int)
    with open(fa, "w") as fh:
        for i, ln in enumerate(lens):
            s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, ln)].tobytes().decode()
            fh.write(f">contig_{i:06d} len={ln}\n")
            fh.write("\n".join(s[k:k + 60] for k in range(0, ln, 60)) + "\n")
    size = fa.stat().st_size
    nw = sequence.ParsedFasta(fa).n_windows
    for rep in range(2):
        t0 = time.perf_counter()
        nn_classification.main(fa, tmp / "out", False, batch, True, 8, False, False)
        dt = time.perf_counter() - t0
        print(f"run {rep}: {size / 1e6:.0f} MB FASTA, {n_contigs} contigs, {nw} windows: {dt:.2f} s wall "
              f"-> {nw / dt:.0f} windows/s, {size / 1e6 / dt:.0f} MB/s", flush=True)
    log = (tmp / "out" / "meta_nn_classification.log").read_text()
    print(log)


if __name__ == "__main__":
    main()
