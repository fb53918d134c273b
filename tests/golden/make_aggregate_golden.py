#!/usr/bin/env python
"""
Golden vectors for aggregated-classification, made by running the REAL reference module in the build container
(numpy only; loaded from /root/reference by file path under a stub ``genomad`` package, like make_golden.py).

  tests/golden/aggregate_golden.npz   -- inputs (w, b1, b2) and the reference ``branch_attention`` outputs (float64),
                                         incl. float32 inputs, extreme marker frequencies and n = 0 / 1.
  tests/golden/aggregate_module/      -- a complete on-disk case: the FASTA, the upstream files the module requires
                                         (written by ``build_case`` below, synthetic), and under ``expected/`` the TSV/NPZ
                                         the reference ``main()`` produced from them (sequence + provirus twins).
Run:  python tests/golden/make_aggregate_golden.py
"""
import importlib.util
import json
import shutil
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/genomad")


def load_reference_module():
    pkg = types.ModuleType("genomad")
    pkg.__path__ = [str(REF)]
    sys.modules["genomad"] = pkg
    for name, rel in (("_paths", "_paths.py"), ("utils", "utils.py"), ("sequence", "sequence.py"),
                      ("aggregated_classification", "modules/aggregated_classification.py")):
        spec = importlib.util.spec_from_file_location(f"genomad.{name}", REF / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"genomad.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    return sys.modules["genomad.aggregated_classification"], sys.modules["genomad.utils"]


def _scores(rng, n, dtype):
    x = rng.random((n, 3)) ** 2 + 1e-3
    return (x / x.sum(1, keepdims=True)).astype(dtype)


def build_case(root: Path, md5_of) -> Path:
    """Write the FASTA and every upstream file aggregated-classification requires; returns the FASTA path."""
    rng = np.random.default_rng(7)
    root.mkdir(parents=True, exist_ok=True)
    fasta = root / "toy.fna"
    names = [f"contig_{i:02d}" for i in range(6)]
    with open(fasta, "w") as f:
        for nm in names:
            f.write(f">{nm} some description\n")
            s = "".join("ACGT"[i] for i in rng.integers(0, 4, 300))
            f.write("\n".join(s[i:i + 60] for i in range(0, 300, 60)) + "\n")
    out = root / "out"
    prov_names = ["contig_01|provirus_10_200", "contig_04|provirus_5_250"]
    info = lambda module: json.dumps({"module": module, "input": fasta.name, "input_md5": md5_of(fasta),
                                      "start_time": "2026-01-01T00:00:00", "parameters": {}}, indent=4) + "\n"
    mk, nn, fp = out / "toy_marker_classification", out / "toy_nn_classification", out / "toy_find_proviruses"
    for d in (mk, nn, fp):
        d.mkdir(parents=True)
    (mk / "toy_marker_classification.json").write_text(info("marker_classification"))
    (nn / "toy_nn_classification.json").write_text(info("nn_classification"))
    (fp / "toy_find_proviruses.json").write_text(info("find_proviruses"))
    feats = rng.random((6, 25)).astype(np.float32)
    feats[2, 15:18] = 0.0                                   # a contig without markers
    pfeats = rng.random((2, 25)).astype(np.float32)
    np.savez_compressed(mk / "toy_features.npz", contig_names=np.array(names), contig_features=feats)
    np.savez_compressed(mk / "toy_provirus_features.npz", provirus_names=np.array(prov_names), provirus_features=pfeats)
    np.savez_compressed(mk / "toy_marker_classification.npz", contig_names=np.array(names),
                        predictions=_scores(rng, 6, np.float32))
    np.savez_compressed(mk / "toy_provirus_marker_classification.npz", provirus_names=np.array(prov_names),
                        predictions=_scores(rng, 2, np.float32))
    np.savez_compressed(nn / "toy_nn_classification.npz", contig_names=np.array(names),
                        predictions=_scores(rng, 6, np.float32))
    np.savez_compressed(nn / "toy_provirus_nn_classification.npz", provirus_names=np.array(prov_names),
                        predictions=_scores(rng, 2, np.float32))
    (fp / "toy_provirus.tsv").write_text("seq_name\tsource_seq\n" + "".join(f"{p}\t{p.split('|')[0]}\n" for p in prov_names))
    (fp / "toy_provirus.fna").write_text("".join(f">{p}\nACGTACGT\n" for p in prov_names))
    (fp / "toy_provirus_proteins.faa").write_text("")
    (fp / "toy_provirus_genes.tsv").write_text("")
    return fasta


def main():
    agg, utils = load_reference_module()
    # the reference prints its installed version in the banner; it is not pip-installed here, so answer that one lookup
    utils.metadata = types.SimpleNamespace(version=lambda _name: "reference-from-source")
    rng = np.random.default_rng(11)
    cases = {}
    for tag, n, dt in (("f32", 257, np.float32), ("f64", 64, np.float64), ("one", 1, np.float32), ("empty", 0, np.float32)):
        w = (rng.random(n) * 3).astype(dt)
        if n > 8:
            w[:4] = [0.0, 1e-6, 3.0, 50.0]                  # no markers ... far outside the trained range
        b1, b2 = _scores(rng, n, dt), _scores(rng, n, dt)
        cases[f"{tag}_w"], cases[f"{tag}_b1"], cases[f"{tag}_b2"] = w, b1, b2
        cases[f"{tag}_out"] = agg.branch_attention(w, b1, b2)
        cases[f"{tag}_out_t1"] = agg.branch_attention(w, b1, b2, temperature=1)
    np.savez_compressed(HERE / "aggregate_golden.npz", **cases)
    print("aggregate_golden.npz written:", sorted(k for k in cases if k.endswith("_out")))

    case_dir = HERE / "aggregate_module"
    if case_dir.exists():
        shutil.rmtree(case_dir)
    with tempfile.TemporaryDirectory() as td:
        fasta = build_case(Path(td), utils.get_md5)
        agg.main(fasta, Path(td) / "out", restart=False, verbose=False)
        produced = Path(td) / "out" / "toy_aggregated_classification"
        (case_dir / "expected").mkdir(parents=True)
        for p in sorted(produced.iterdir()):
            if p.suffix in (".tsv", ".npz"):
                shutil.copy(p, case_dir / "expected" / p.name)
        print("reference main() outputs:", sorted(p.name for p in produced.iterdir()))


if __name__ == "__main__":
    main()
