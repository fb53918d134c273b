"""CPU tests: the oracle against golden vectors made by the real reference code (tests/golden/make_golden.py)."""
import hashlib
import json

import numpy as np
import pytest
import torch

from oracle import igloo_model as M
from oracle import tokenizer as T


@pytest.fixture(scope="module")
def enc(golden_dir):
    return json.loads((golden_dir / "encoder_golden.json").read_text())


def test_tokenizer_literal_matches_reference(enc):
    for case in enc["tokenize"]:
        assert T.tokenize_dna_literal(case["seq"].encode("ascii")) == case["tokens"], case["seq"]


def test_survey_kat():
    assert T.tokenize_dna_literal(b"ACGTNACGTACGTTTTT") == [28, 0, 0, 0, 0, 28, 109, 178, 199, 28, 112, 192, 256, 256]


def test_tokenizer_closed_form_matches_reference_batch(enc, golden_dir):
    ref = np.load(golden_dir / "encoder_batch_tokens.npz")["tokens"]
    raws = enc["batch"]["raw"]
    a = np.frombuffer(b"".join(r.upper().encode("ascii").ljust(6000, b"N") for r in raws), np.uint8).reshape(-1, 6000)
    tok = T.tokenize_windows(a)
    assert tok.dtype == np.uint16 and tok.shape == (len(raws), 5997)
    assert np.array_equal(tok, ref)
    assert hashlib.sha256(tok.tobytes()).hexdigest() == enc["batch"]["tokens_sha256"]
    # literal loop agrees too (a few rows; it is slow)
    for i in (0, 1, 5):
        assert T.tokenize_dna_literal(a[i].tobytes()) == ref[i].tolist()


def test_seq_windows_lengths(enc):
    for case in enc["windows"]:
        s = "A" * case["len"]
        assert [len(w) for w in T.seq_windows(s)] == case["multi"]
        assert [len(w) for w in T.seq_windows(s, max_windows=1)] == case["single"]


def test_n_rule(enc, tmp_path):
    for case in enc["nrule"]:
        p = tmp_path / "x.fna"
        p.write_text(f">x\n{case['raw']}\n")
        names, ids, ascii_arr, tok = T.encode_fasta(p)
        assert len(ids) == len(case["kept"]), case["name"]


def test_read_fasta_quirks(enc, tmp_path):
    for name, case in enc["fasta"].items():
        p = tmp_path / f"{name}.fna"
        p.write_text(case["text"], newline="")
        got = [[h, T.accession(h), s] for h, s in T.read_fasta(p)]
        assert got == [list(r) for r in case["records"]], name


def test_model_frozen_vectors(golden_dir, weights_npz):
    g = np.load(golden_dir / "model_golden.npz")
    w = M.load_npz_weights(weights_npz)
    tok = T.tokenize_windows(g["ascii"])
    p = M.forward(tok[:6], w, torch.float64)
    assert np.abs(p - g["shipped_fp64"][:6]).max() < 1e-12
    # SURVEY Appendix C cross-check vectors (fp64), independently derived during the survey
    appendix_c = np.array([[0.00000000, 0.00011596, 0.99988404], [0.99973547, 0.00026453, 0.0],
                           [0.19144470, 0.05451930, 0.75403600], [0.96907142, 0.03092858, 0.0],
                           [0.05019702, 0.06085832, 0.88894466]])
    assert np.abs(p[:5] - appendix_c).max() < 5e-9


def test_two_formulations_agree(golden_dir, weights_npz):
    g = np.load(golden_dir / "model_golden.npz")
    w = M.load_npz_weights(weights_npz)
    tok = T.tokenize_windows(g["ascii"][2:6])
    for ww in (w, M.synthetic_igloo_weights(w)):
        a = M.forward(tok, ww, torch.float64)
        b = M.forward_as_written(tok, ww, torch.float64)
        assert np.abs(a - b).max() < 1e-12
        c = M.forward_as_written(tok, ww, torch.float32)
        assert np.abs(a - c).max() < 5e-5


def test_synthetic_weights_make_attention_live(golden_dir, weights_npz):
    g = np.load(golden_dir / "model_golden.npz")
    w = M.load_npz_weights(weights_npz)
    ws = M.synthetic_igloo_weights(w)
    tok = T.tokenize_windows(g["ascii"][4:6])
    _, parts = M.forward(tok, w, torch.float32, return_intermediates=True)
    assert float(parts["ig0"]["logits"].abs().max()) == 0.0  # shipped weights: attention is exactly uniform
    _, parts = M.forward(tok, ws, torch.float64, return_intermediates=True)
    assert float(parts["ig0"]["alpha"].max()) > 5.0 / 749    # synthetic weights: far from uniform


def test_segment_mean():
    p = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], np.float32)
    out = T.segment_mean(p, np.array([0, 0, 2, 2]))
    assert out.shape == (3, 3)
    assert np.allclose(out, [[.5, .5, 0], [0, 0, 0], [.5, .5, 1]])
