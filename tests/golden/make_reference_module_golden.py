#!/usr/bin/env python
"""
Golden OUTPUT DIRECTORY of the reference's nn-classification module: /root/reference/genomad/modules/nn_classification.py::main is
executed unmodified -- with the reference's own sequence.py (numba), utils.py, _paths.py, neural_network/{model,igloo}.py and
nn_classifier.h5 -- on a small FASTA with a finished find-proviruses run next to it.  Only `tensorflow` and `keras` are replaced, by
the NumPy stand-in tests/golden/keras_shim.py (neither library exists in this image).  Two runs: default, and --single-window
--cleanup into a second directory.

    python tests/golden/make_reference_module_golden.py    # needs /root/reference; ~1 min -> tests/golden/reference_module/

What is stored: the input FASTA, the find-proviruses files, and for each run the produced TSVs, NPZ contents (as .npz), the
execution-info JSON, the tokens of the encoding stage (read back from the .tfrec files of the default run), the list of files the module left behind, and the log with timestamps stripped.  The tests compare
genomad_b200's module against these (names, NPZ keys / dtypes, TSV bytes up to the 4th decimal of the scores, JSON keys, file set).
"""
import importlib.util
import json
import re
import shutil
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference/genomad")
HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))
import keras_shim  # noqa: E402


def load_reference():
    keras_shim.install()
    pkg = types.ModuleType("genomad")
    pkg.__path__ = [str(REF)]
    sys.modules["genomad"] = pkg

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(f"genomad.{name}", REF / rel)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"genomad.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name.split(".")[-1], mod)
        return mod
    load("_paths", "_paths.py")
    utils = load("utils", "utils.py")
    # the checkout is not pip-installed: importlib.metadata has no "genomad" distribution (only the version in the header panel)
    version = re.search(r'^__version__\s*=\s*"([^"]+)"', (REF / "__init__.py").read_text(), re.M).group(1)
    utils.metadata = types.SimpleNamespace(version=lambda name: version)
    load("sequence", "sequence.py")
    nn = types.ModuleType("genomad.neural_network")
    nn.__path__ = [str(REF / "neural_network")]
    sys.modules["genomad.neural_network"] = nn
    pkg.neural_network = nn
    for name in ("igloo", "model"):
        mod = load(f"neural_network.{name}", f"neural_network/{name}.py")
        setattr(nn, name, mod)
    nn.create_classifier = sys.modules["genomad.neural_network.model"].create_classifier
    return load("nn_classification", "modules/nn_classification.py"), sys.modules["genomad.utils"], sys.modules["genomad._paths"]


def make_inputs(root: Path, utils, paths):
    rng = np.random.default_rng(11)

    def seq(n, alphabet=b"ACGT"):
        return np.frombuffer(alphabet, np.uint8)[rng.integers(0, len(alphabet), n)].tobytes().decode()
    root.mkdir(parents=True)
    fa = root / "toy.fna"
    s3 = seq(6000) + "N" * 4500 + seq(1500) + seq(3000)          # 2nd window: 4,500 N > 4,000 -> skipped; 3rd window kept
    records = [("ctg_a desc 1", seq(14500)),                     # 3 windows (6000, 6000, 2500)
               ("ctg_b", seq(2500).lower()),                     # lower case, one short window
               ("ctg_c x", s3),
               ("ctg_d", seq(8400, b"ACGTRYN")),                 # IUPAC / N scattered; 2nd window 2,400 < 2,500 -> dropped
               ("ctg_e", seq(6000))]
    with open(fa, "w") as fh:
        for i, (name, s) in enumerate(records):
            width = (60, 80, 70, 61, 100)[i]
            fh.write(f">{name}\n" + "\n".join(s[k:k + width] for k in range(0, len(s), width)) + "\n")
    out = root / "out"
    out.mkdir()
    o = paths.GenomadOutputs("toy", out)
    o.find_proviruses_dir.mkdir()
    utils.write_execution_info("find_proviruses", fa, {}, o.find_proviruses_execution_info)
    pv = {"ctg_a|provirus_1001_13000": seq(12000), "ctg_c|provirus_1_3000": seq(3000)}
    o.find_proviruses_output.write_text("seq_name\tx\n" + "".join(f"{k}\t1\n" for k in pv))
    o.find_proviruses_nucleotide_output.write_text("".join(f">{k}\n{v}\n" for k, v in pv.items()))
    o.find_proviruses_proteins_output.write_text("")
    o.find_proviruses_genes_output.write_text("")
    return fa, out


def snapshot(out: Path, dst: Path, paths, prefix="toy"):
    o = paths.GenomadOutputs(prefix, out)
    dst.mkdir(parents=True)
    files = sorted(str(p.relative_to(out)) for p in out.rglob("*") if p.is_file() and "find_proviruses" not in str(p))
    (dst / "files.json").write_text(json.dumps(files, indent=1) + "\n")
    for p in (o.nn_classification_output, o.provirus_nn_classification_output, o.nn_classification_execution_info):
        if p.exists():
            shutil.copy(p, dst / p.name)
    for p in (o.nn_classification_npz_output, o.provirus_nn_classification_npz_output, o.seq_window_id_output,
              o.provirus_window_id_output):
        if p.exists():
            shutil.copy(p, dst / p.name)
    log = o.nn_classification_log.read_text()
    (dst / "log_without_timestamps.txt").write_text(re.sub(r"^\[\d\d:\d\d:\d\d\] ", "", log, flags=re.M))


def main():
    np.random.seed(0)
    mod, utils, paths = load_reference()
    gold = HERE / "reference_module"
    if gold.exists():
        shutil.rmtree(gold)
    work = Path("/tmp/reference_module_work")
    if work.exists():
        shutil.rmtree(work)
    fa, out = make_inputs(work, utils, paths)
    (gold / "input").mkdir(parents=True)
    shutil.copy(fa, gold / "input" / fa.name)
    shutil.copytree(out / "toy_find_proviruses", gold / "input" / "toy_find_proviruses")
    # run 1: defaults (batch size 4 so that several batches and a ragged last one occur)
    mod.main(fa, out, False, 4, False, 1, False, False)
    snapshot(out, gold / "run_default", paths)
    # the tokens the reference's encoding stage produced (sequence.tokenize_dna under numba), read back from its .tfrec files
    tf = sys.modules["tensorflow"]
    desc = {"sequence": tf.io.FixedLenFeature([5997], tf.int64)}
    toks = {}
    for key, sub in (("sequences", "toy_encoded_sequences"), ("proviruses", "toy_encoded_proviruses")):
        files = utils.natsort(tf.io.gfile.glob(f"{out}/toy_nn_classification/{sub}/*.tfrec"))
        rows = [tf.io.parse_single_example(r, desc)["sequence"] for r in tf.data.TFRecordDataset(files)]
        toks[key] = np.stack(rows).astype(np.uint16)
    np.savez_compressed(gold / "run_default" / "encoded_tokens.npz", **toks)
    # skip / restart scenarios in the same directory (log messages only): nothing changed; provirus NPZ lost; --restart;
    # parameter change (--single-window, with --cleanup)
    o = paths.GenomadOutputs("toy", out)
    scen = {}

    def logged(name, *args):
        mod.main(fa, out, *args)
        scen[name] = re.sub(r"^\[\d\d:\d\d:\d\d\] ", "", o.nn_classification_log.read_text(), flags=re.M)
    logged("rerun_unchanged", False, 4, False, 1, False, False)
    o.provirus_nn_classification_npz_output.unlink()
    logged("provirus_npz_lost", False, 4, False, 1, False, False)
    logged("restart", False, 4, True, 1, False, False)
    logged("parameter_change_single_window_cleanup", True, 4, False, 1, False, True)
    (gold / "scenario_logs.json").write_text(json.dumps(scen, indent=1) + "\n")
    # run 2: --single-window --cleanup into a fresh directory (same find-proviruses files)
    out2 = work / "out_single"
    shutil.copytree(gold / "input" / "toy_find_proviruses", out2 / "toy_find_proviruses")
    mod.main(fa, out2, True, 4, False, 1, False, True)
    snapshot(out2, gold / "run_single_window_cleanup", paths)
    print((gold / "run_default" / "toy_nn_classification.tsv").read_text())
    print((gold / "run_default" / "files.json").read_text())
    print((gold / "run_default" / "log_without_timestamps.txt").read_text())


if __name__ == "__main__":
    main()
