#!/usr/bin/env python
"""
Generate golden vectors by running the REAL reference code in the build container.

Only the tokenizer / windowing / FASTA reader can run here (numba is installed; TensorFlow,
Keras and h5py are not), so this produces:

  tests/golden/encoder_golden.json   -- reference ``tokenize_dna`` outputs on adversarial strings,
                                        ``seq_windows`` lengths + the N>4000 rule, ``read_fasta``
                                        (strip_n=True) on quirky FASTA text, plus a sha256 of the
                                        reference tokens of a seeded 64x6000 batch.
  tests/golden/model_golden.npz      -- frozen outputs of the *oracle* (fp64 and fp32) on fixed
                                        windows with the shipped and the synthetic weights. These
                                        are NOT reference outputs (parity unpinned, see oracle/__init__.py);
                                        they freeze the restatement so regressions are caught.

The reference is imported from /root/reference by file path under a stub ``genomad`` package
(``import genomad`` itself needs xgboost/pycrfsuite/taxopy, which are absent).
Run:  python tests/golden/make_golden.py
"""
import hashlib
import importlib.util
import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference/genomad")
sys.path.insert(0, str(ROOT))


def load_reference_sequence():
    pkg = types.ModuleType("genomad")
    pkg.__path__ = [str(REF)]
    sys.modules["genomad"] = pkg
    for name in ("_paths", "utils", "sequence"):
        spec = importlib.util.spec_from_file_location(f"genomad.{name}", REF / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"genomad.{name}"] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    return sys.modules["genomad.sequence"]


ADVERSARIAL = [
    "ACGTNACGTACGTTTTT", "ACGT", "ACG", "A", "", "NNNN", "NACGT", "ACGTN", "ANCGTACGT", "ACGNTACGT",
    "ACGTRYKMACGTACGT", "acgtacgtacgt", "ACGT-ACGT*ACGT", "ACGTACGT\rACGT", "TTTTTTTT", "AAAAAAAA",
    "ACGTACGTNNACGTACGTNACGTNNNACGTACGT", "NNNACGTACGTNNN", "ACGUACGU", "ACGT ACGT", "GATTACA" * 9,
    "N" * 17 + "ACGTACGT", "ACGTACGT" + "N" * 17, "ANANANANANACGTACGTAN",
]

FASTA_CASES = {
    "plain": ">c1 desc here\nACGTACGT\nACGT\n>c2\nNNNNACGTNN\n>c3\nNNNN\n>c4\tt\nacgtnn\n",
    "leading_text": "junk line\nmore junk\n>c1\nACGT\n\n>c2\nAC\n\nGT\n",
    "crlf": ">c1 x\r\nACGT\r\nACGT\r\n>c2\r\nTTTT\r\n",
    "no_trailing_newline": ">c1\nACGTNN\n>c2\nnnACGTnn",
    "empty_records": ">c1\n>c2\nACGT\n>c3\n\n>c4\nGG\n",
    "inner_n": ">c1\nNNACNNGTNN\n",
}

WINDOW_LENGTHS = [1, 400, 2499, 2500, 5999, 6000, 6001, 8400, 8499, 8500, 12000, 14000, 14499, 14500, 20000]


def main():
    seq = load_reference_sequence()
    out = {"tokenize": [], "windows": [], "fasta": {}, "nrule": []}
    for s in ADVERSARIAL:
        out["tokenize"].append({"seq": s, "tokens": [int(x) for x in seq.tokenize_dna(s.encode("ascii"), 4)]})
    # padded-window form, exactly what generate_data does (nn_classification.py:72-73)
    rng = np.random.default_rng(123)
    batch = []
    alphabet = np.frombuffer(b"ACGTNRYacgtn-", np.uint8)
    probs = np.array([.22, .22, .22, .22, .04, .01, .01, .01, .01, .01, .01, .01, .01])
    for i in range(64):
        ln = int(rng.integers(1, 6001)) if i % 4 == 0 else 6000
        raw = alphabet[rng.choice(len(alphabet), ln, p=probs)].tobytes().decode()
        s = seq.Sequence("x", raw)
        padded = s.seq_ascii.ljust(6000, b"N")
        toks = np.array(seq.tokenize_dna(padded, 4), dtype=np.uint16)
        assert toks.shape == (5997,)
        batch.append((raw, toks))
    h = hashlib.sha256()
    for raw, toks in batch:
        h.update(toks.tobytes())
    out["batch"] = {"seed": 123, "raw": [b[0] for b in batch], "tokens_sha256": h.hexdigest(),
                    "first_tokens": [[int(x) for x in b[1][:12]] for b in batch]}
    np.savez_compressed(Path(__file__).parent / "encoder_batch_tokens.npz",
                        tokens=np.stack([b[1] for b in batch]))
    for ln in WINDOW_LENGTHS:
        s = seq.Sequence("x", "A" * ln)
        out["windows"].append({
            "len": ln,
            "multi": [len(wd) for wd in seq.seq_windows(s, 6000, 2500, max_windows=None)],
            "single": [len(wd) for wd in seq.seq_windows(s, 6000, 2500, max_windows=1)],
        })
    # N rule (nn_classification.py:70-71): literal "N" count on the raw window, first window exempt
    for name, raw in {
        "second_window_4001_N": "A" * 6000 + "N" * 4001 + "C" * 1999,
        "second_window_4000_N": "A" * 6000 + "N" * 4000 + "C" * 2000,
        "second_window_lower_n": "A" * 6000 + "n" * 5000 + "C" * 1000,
        "first_window_all_N_inner": "A" + "N" * 5998 + "C" + "G" * 3000,
    }.items():
        s = seq.Sequence("x", raw)
        kept = []
        for window_n, wd in enumerate(seq.seq_windows(s, 6000, 2500)):
            if window_n > 0 and wd.count("N") > 4000:
                continue
            kept.append(window_n)
        out["nrule"].append({"name": name, "raw": raw, "kept": kept})
    with tempfile.TemporaryDirectory() as td:
        for name, text in FASTA_CASES.items():
            p = Path(td) / f"{name}.fna"
            p.write_text(text, newline="")
            recs = [(r.header, r.accession, r.seq) for r in seq.read_fasta(p, strip_n=True)]
            out["fasta"][name] = {"text": text, "records": recs, "check_fasta": bool(seq.check_fasta(p))}
    (Path(__file__).parent / "encoder_golden.json").write_text(json.dumps(out, indent=1))
    print("encoder_golden.json written;", len(out["tokenize"]), "tokenize cases")

    # ---- frozen oracle outputs (NOT reference outputs)
    import torch
    from oracle import igloo_model as M, tokenizer as T

    def lcg():
        x, o = 42, []
        for _ in range(6000):
            x = (1103515245 * x + 12345) % (2 ** 31)
            o.append("ACGT"[(x >> 16) & 3])
        return "".join(o).encode()
    wins = [b"A" * 6000, b"ACGT" * 1500, b"N" * 6000, (b"ACGT" * 625).ljust(6000, b"N"), lcg()]
    r = np.random.default_rng(5)
    for _ in range(11):
        wins.append(np.frombuffer(b"ACGT", np.uint8)[r.integers(0, 4, 6000)].tobytes())
    a = np.frombuffer(b"".join(wins), np.uint8).reshape(-1, 6000)
    tok = T.tokenize_windows(a)
    w = M.load_npz_weights(ROOT / "genomad_b200/data/nn_classifier.npz")
    ws = M.synthetic_igloo_weights(w)
    np.savez_compressed(
        Path(__file__).parent / "model_golden.npz", ascii=a,
        shipped_fp64=M.forward(tok, w, torch.float64), shipped_fp32=M.forward(tok, w, torch.float32),
        synthetic_fp64=M.forward(tok, ws, torch.float64), synthetic_fp32=M.forward(tok, ws, torch.float32))
    print("model_golden.npz written")


if __name__ == "__main__":
    main()
