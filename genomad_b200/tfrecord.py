"""
TFRecord files of tokenised windows -- the reference's ``<prefix>_encoded_sequences/<count>.tfrec`` intermediates
(reference nn_classification.py:43-52 writer, :87-100 reader).  Thin ctypes wrappers around the
native writer/reader in libgnm.so (csrc/tfrecord.cpp); off by default in the module driver because nothing reads them.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import numpy as np

from . import engine

TOKENS = 5997
RECORDS_PER_FILE = 10_000          # reference generate_data(n_records_per_file=10_000), nn_classification.py:58


def _threads(threads: Optional[int]) -> int:
    return int(threads or min(16, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 4))


def crc32c(data: bytes) -> int:
    lib = engine.load_library()
    return int(lib.gnm_crc32c(C.cast(C.c_char_p(data), C.c_void_p), len(data)))


def write_tfrecord(path, tokens: np.ndarray, threads: Optional[int] = None) -> None:
    """tokens: uint16 [n, 5997] -> one TFRecord file (one Example{"sequence": Int64List} per row)."""
    lib = engine.load_library()
    t = np.ascontiguousarray(tokens, dtype=np.uint16)
    if t.ndim != 2 or t.shape[1] != TOKENS:
        raise ValueError(f"tokens must be [n, {TOKENS}], got {t.shape}")
    if lib.gnm_tfrecord_write(str(path).encode(), t.ctypes.data if len(t) else None, len(t), _threads(threads)) != 0:
        raise RuntimeError(lib.gnm_tfrecord_last_error().decode())


def count_records(path) -> int:
    """Number of records (both CRCs of every record are verified on the way)."""
    lib = engine.load_library()
    n = C.c_int64()
    if lib.gnm_tfrecord_read(str(path).encode(), None, 0, C.byref(n)) != 0:
        raise RuntimeError(lib.gnm_tfrecord_last_error().decode())
    return n.value


def read_tfrecord(path) -> np.ndarray:
    """One TFRecord file -> uint16 [n, 5997]; raises on CRC mismatch or any record that is not the layout above."""
    lib = engine.load_library()
    n = count_records(path)
    out = np.empty((n, TOKENS), dtype=np.uint16)
    got = C.c_int64()
    if lib.gnm_tfrecord_read(str(path).encode(), out.ctypes.data if n else None, n, C.byref(got)) != 0:
        raise RuntimeError(lib.gnm_tfrecord_last_error().decode())
    return out[: got.value]


def tfrecord_files(directory) -> list:
    """The directory's .tfrec files in the order the reference reads them (numeric stem = cumulative window count;
    reference nn_classification.py:293-295 natsorts the glob)."""
    return sorted(Path(directory).glob("*.tfrec"), key=lambda p: int(p.stem))
