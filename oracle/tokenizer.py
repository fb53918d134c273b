"""
CPU restatement of the reference's FASTA reader, windowing and 4-mer tokenizer.
TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED against the real reference
code through tests/golden/encoder_golden.json (made by tests/golden/make_golden.py).

Follows, line by line:
  * ``tokenize_dna``          reference genomad/sequence.py:170-193
  * ``seq_windows``           reference genomad/sequence.py:150-167
  * ``read_fasta``            reference genomad/sequence.py:96-121 (strip_n=True path)
  * window rules / padding    reference genomad/modules/nn_classification.py:65-73
"""
from __future__ import annotations

import bz2
import gzip
import lzma
from typing import Iterator, List, Optional, Tuple

import numpy as np

WINDOW = 6000
MIN_TAIL = 2500
MAX_N = 4000
WORD = 4
NTOK = WINDOW - WORD + 1  # 5997


def tokenize_dna_literal(seq: bytes, word_size: int = WORD) -> List[int]:
    """Pure-Python transcription of the numba loop (sequence.py:170-193). Small inputs only."""
    final_length = len(seq) - word_size + 1
    out: List[int] = []
    kmer = 0
    countdown = word_size - 1
    mask = (1 << 2 * word_size) - 1
    for base in seq:
        if base == 65:
            kmer = ((kmer << 2) | 0) & mask
        elif base == 67:
            kmer = ((kmer << 2) | 1) & mask
        elif base == 71:
            kmer = ((kmer << 2) | 2) & mask
        elif base == 84:
            kmer = ((kmer << 2) | 3) & mask
        else:
            out += [0] * (word_size - countdown)
            countdown = word_size
        if countdown == 0:
            out.append(kmer + 1)
        else:
            countdown -= 1
    # Python slicing semantics: a negative final_length drops from the end
    return out[:final_length]


_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i


def tokenize_windows(ascii_windows: np.ndarray) -> np.ndarray:
    """
    Vectorised closed form of ``tokenize_dna(.., 4)`` for a batch of equal-length windows.

    ascii_windows: uint8 [n, L] (already upper-cased and N-padded by the caller, as
    nn_classification.py:72 does).  Returns uint16 [n, L-3]; token = 0 if any of the four
    bytes is not one of ``A C G T`` (uppercase ASCII), else 1 + base-4 value of the 4-mer.
    """
    a = np.asarray(ascii_windows, dtype=np.uint8)
    c = _CODE[a]
    bad = c == 255
    c = np.where(bad, 0, c).astype(np.uint16)
    tok = 1 + 64 * c[:, :-3] + 16 * c[:, 1:-2] + 4 * c[:, 2:-1] + c[:, 3:]
    anybad = bad[:, :-3] | bad[:, 1:-2] | bad[:, 2:-1] | bad[:, 3:]
    return np.where(anybad, 0, tok).astype(np.uint16)


def _open_text(path):
    with open(path, "rb") as fh:
        sig = fh.read(8)
    if sig[:2] == b"\x1f\x8b":
        return gzip.open(path, "rt")
    if sig[:3] == b"BZh":
        return bz2.open(path, "rt")
    if sig[:7] == b"\xfd7zXZ\x00\x00":
        return lzma.open(path, "rt")
    return open(path, "r")


def read_fasta(path, strip_n: bool = True) -> Iterator[Tuple[str, str]]:
    """Yield (header, sequence) exactly as sequence.py:96-121 would (records empty after strip are dropped)."""
    with _open_text(path) as fin:
        last = None
        while True:
            if not last:
                for line in fin:
                    if line[0] == ">":
                        last = line.removesuffix("\n")
                        break
            if not last:
                break
            name, seqs, last = last[1:], [], None
            for line in fin:
                if line[0] == ">":
                    last = line.removesuffix("\n")
                    break
                seqs.append(line.removesuffix("\n"))
            seq = "".join(seqs)
            if strip_n:
                seq = seq.strip("nN")
            if len(seq):
                yield name, seq
            if not last:
                break


def accession(header: str) -> str:
    """sequence.py:24-25"""
    return header.split()[0]


def seq_windows(seq: str, length: int = WINDOW, min_length: int = MIN_TAIL,
                force_first_window: bool = True, max_windows: Optional[int] = None) -> Iterator[str]:
    """sequence.py:150-167 on a plain str."""
    win = 0
    while win * length < len(seq):
        w = seq[win * length:(win + 1) * length]
        if len(w) < min_length:
            if win == 0 and force_first_window:
                yield w
            break
        yield w
        win += 1
        if max_windows and win == max_windows:
            break


def encode_fasta(path, single_window: bool = False):
    """
    generate_data() of nn_classification.py:54-82 without the TFRecord round trip.

    Returns (contig_names [n] str array, contig_ids int64 [W], ascii uint8 [W,6000], tokens uint16 [W,5997]).
    """
    names: List[str] = []
    ids: List[int] = []
    wins: List[bytes] = []
    max_windows = 1 if single_window else None
    for contig_id, (header, seq) in enumerate(read_fasta(path, strip_n=True)):
        names.append(accession(header))
        for window_n, w in enumerate(seq_windows(seq, WINDOW, MIN_TAIL, max_windows=max_windows)):
            if window_n > 0 and w.count("N") > MAX_N:
                continue
            wins.append(w.upper().encode("ascii").ljust(WINDOW, b"N"))
            ids.append(contig_id)
    ascii_arr = (np.frombuffer(b"".join(wins), dtype=np.uint8).reshape(-1, WINDOW)
                 if wins else np.zeros((0, WINDOW), np.uint8))
    return (np.array(names), np.array(ids, dtype=np.int64), ascii_arr, tokenize_windows(ascii_arr))


def segment_mean(preds: np.ndarray, ids: np.ndarray, n_segments: Optional[int] = None) -> np.ndarray:
    """
    tf.math.segment_mean(preds, ids) (nn_classification.py:320): fp32 running sum in row
    order, divided by the count.  ids must be sorted; absent ids give zero rows.
    """
    preds = np.asarray(preds, dtype=np.float32)
    ids = np.asarray(ids, dtype=np.int64)
    n = int(ids[-1]) + 1 if n_segments is None and len(ids) else (n_segments or 0)
    s = np.zeros((n, preds.shape[1]), dtype=np.float32)
    c = np.zeros(n, dtype=np.float32)
    np.add.at(s, ids, preds)  # unbuffered, sequential in row order
    np.add.at(c, ids, np.float32(1))
    return (s / np.maximum(c, 1)[:, None]).astype(np.float32)
