"""
aggregated-classification against golden vectors produced by the REAL reference module
(tests/golden/make_aggregate_golden.py): the combiner is float64 NumPy in both, so the bar is bit-exact.
"""
import importlib.util
import shutil
from pathlib import Path

import numpy as np
import pytest

from genomad_b200 import aggregated_classification as agg, utils

GOLD = Path(__file__).parent / "golden"


def _generator():
    spec = importlib.util.spec_from_file_location("make_aggregate_golden", GOLD / "make_aggregate_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # importing it does not touch /root/reference (only its main() does)
    return mod


@pytest.mark.parametrize("tag", ["f32", "f64", "one", "empty"])
def test_branch_attention_bit_exact(tag):
    z = np.load(GOLD / "aggregate_golden.npz")
    w, b1, b2 = z[f"{tag}_w"], z[f"{tag}_b1"], z[f"{tag}_b2"]
    for key, kwargs in ((f"{tag}_out", {}), (f"{tag}_out_t1", {"temperature": 1})):
        got = agg.branch_attention(w, b1, b2, **kwargs)
        assert got.dtype == np.float64 and got.shape == z[key].shape
        assert np.array_equal(got, z[key])
    if len(w):
        got = agg.branch_attention(w, b1, b2)
        assert np.allclose(got.sum(1), 1.0, atol=1e-12)


def test_aggregate_in_memory_matches_file_path():
    rng = np.random.default_rng(0)
    feats = rng.random((9, 25)).astype(np.float32)
    b1, b2 = rng.random((9, 3)).astype(np.float32), rng.random((9, 3)).astype(np.float32)
    assert np.array_equal(agg.aggregate_in_memory(feats, b1, b2),
                          agg.branch_attention(feats[:, 15:18].sum(1), b1, b2))


@pytest.fixture
def case(tmp_path):
    fasta = _generator().build_case(tmp_path, utils.get_md5)
    return fasta, tmp_path / "out"


def _check_against_reference(out_dir):
    exp = GOLD / "aggregate_module" / "expected"
    got = out_dir / "toy_aggregated_classification"
    for name in ("toy_aggregated_classification.tsv", "toy_provirus_aggregated_classification.tsv"):
        assert (got / name).read_bytes() == (exp / name).read_bytes(), name
    for name, key in (("toy_aggregated_classification.npz", "contig_names"),
                      ("toy_provirus_aggregated_classification.npz", "provirus_names")):
        a, b = np.load(got / name), np.load(exp / name)
        assert sorted(a.files) == sorted(b.files)
        assert a["predictions"].dtype == b["predictions"].dtype == np.float64
        assert np.array_equal(a["predictions"], b["predictions"]) and list(a[key]) == list(b[key])


def test_module_matches_reference_outputs(case):
    fasta, out = case
    agg.main(fasta, out, restart=False, verbose=False)
    _check_against_reference(out)
    info = utils.get_execution_info(out / "toy_aggregated_classification" / "toy_aggregated_classification.json")
    assert info[0] == utils.get_md5(fasta) and info[1] == "aggregated_classification" and info[2] == {}
    log = (out / "toy_aggregated_classification.log").read_text()
    assert "Sequences classified." in log and "Proviruses classified." in log and "finished!" in log


def test_skip_and_restart(case):
    fasta, out = case
    agg.main(fasta, out, restart=False, verbose=False)
    nn_npz = out / "toy_nn_classification" / "toy_nn_classification.npz"
    z = np.load(nn_npz)
    np.savez_compressed(nn_npz, contig_names=z["contig_names"], predictions=z["predictions"][::-1].copy())
    agg.main(fasta, out, restart=False, verbose=False)                 # same input md5 -> NPZ reused, TSV rewritten
    log = (out / "toy_aggregated_classification.log").read_text()
    assert "toy_aggregated_classification.npz was found. Skipping sequence classification." in log
    assert "toy_provirus_aggregated_classification.npz was found. Skipping provirus classification." in log
    _check_against_reference(out)
    agg.main(fasta, out, restart=True, verbose=False)
    new = np.load(out / "toy_aggregated_classification" / "toy_aggregated_classification.npz")["predictions"]
    old = np.load(GOLD / "aggregate_module" / "expected" / "toy_aggregated_classification.npz")["predictions"]
    assert not np.array_equal(new, old)


def test_without_proviruses(case):
    fasta, out = case
    shutil.rmtree(out / "toy_find_proviruses")
    agg.main(fasta, out, restart=False, verbose=False)
    d = out / "toy_aggregated_classification"
    assert (d / "toy_aggregated_classification.tsv").exists()
    assert not (d / "toy_provirus_aggregated_classification.tsv").exists()


def test_errors(case, capsys):
    fasta, out = case
    (out / "toy_nn_classification" / "toy_nn_classification.npz").unlink()
    with pytest.raises(SystemExit) as e:
        agg.main(fasta, out, restart=False, verbose=False)
    assert e.value.code == 1
    assert "toy_nn_classification.npz" in capsys.readouterr().err
    assert not (out / "toy_aggregated_classification").exists()


def test_md5_mismatch(case, capsys):
    fasta, out = case
    other = fasta.with_name("other.fna")
    other.write_text(">x\nACGT\n")
    for mod in ("marker_classification", "nn_classification"):
        src = out / f"toy_{mod}"
        dst = out / f"other_{mod}"
        shutil.copytree(src, dst)
        for p in list(dst.iterdir()):
            p.rename(dst / p.name.replace("toy_", "other_", 1))
    with pytest.raises(SystemExit) as e:
        agg.main(other, out, restart=False, verbose=False)
    assert e.value.code == 1
    assert "Different input FASTA files" in capsys.readouterr().err


def test_cli_surface():
    from click.testing import CliRunner
    from genomad_b200.cli import cli
    r = CliRunner().invoke(cli, ["aggregated-classification", "--help"])
    assert r.exit_code == 0 and "--restart" in r.output and "--quiet" in r.output
