"""
TEST INFRASTRUCTURE ONLY.  Compiles oracle/tokenizer_c.c (the plain-C restatement of the tokenizer and the window rules) into
oracle/_build/liboracle_tok.so with gcc and binds it through ctypes.  Called by __graft_entry__.build() and by the tests.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "tokenizer_c.c"
LIB = HERE / "_build" / "liboracle_tok.so"
_lib = None


def ensure_built(force: bool = False) -> Path:
    """(Re)build when the source's sha256 differs from the one recorded next to the library (file times do not survive a copy of
    the tree).  If gcc is missing but a library exists, the existing one is kept."""
    import hashlib
    digest = hashlib.sha256(SRC.read_bytes()).hexdigest()
    stamp = LIB.with_suffix(".so.src_sha256")
    if not force and LIB.exists() and stamp.exists() and stamp.read_text().strip() == digest:
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    try:
        subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-shared", "-fPIC", str(SRC), "-o", str(LIB)], check=True)
        stamp.write_text(digest + "\n")
    except (OSError, subprocess.CalledProcessError):
        if not LIB.exists():
            raise
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(ensure_built()))
        lib.gnm_oracle_tokenize.restype = C.c_int64
        lib.gnm_oracle_tokenize.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        lib.gnm_oracle_tokenize_windows.restype = C.c_int64
        lib.gnm_oracle_tokenize_windows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        lib.gnm_oracle_window_plan.restype = C.c_int64
        lib.gnm_oracle_window_plan.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        _lib = lib
    return _lib


def tokenize(seq: bytes) -> list:
    """tokenize_dna(seq, 4) for any length."""
    a = np.frombuffer(seq, np.uint8) if len(seq) else np.zeros(0, np.uint8)
    a = np.ascontiguousarray(a)
    out = np.zeros(len(seq) + 4, np.uint16)
    n = load().gnm_oracle_tokenize(a.ctypes.data, len(seq), out.ctypes.data)
    return out[:n].tolist()


def tokenize_windows(ascii_windows: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(ascii_windows, np.uint8)
    n, length = a.shape
    out = np.empty((n, length - 3), np.uint16)
    scratch = np.empty(length + 4, np.uint16)
    rc = load().gnm_oracle_tokenize_windows(a.ctypes.data, n, length, out.ctypes.data, scratch.ctypes.data)
    if rc != 0:
        raise ValueError(f"window {-rc - 1}: token count != length - 3")
    return out


def window_plan(seq: bytes, single_window: bool = False):
    """-> (starts, lengths) of the windows of one stripped record that are classified."""
    a = np.ascontiguousarray(np.frombuffer(seq, np.uint8))
    cap = len(seq) // 6000 + 2
    s, l = np.empty(cap, np.int64), np.empty(cap, np.int64)
    n = load().gnm_oracle_window_plan(a.ctypes.data, len(seq), int(single_window), s.ctypes.data, l.ctypes.data, cap)
    assert n >= 0
    return s[:n], l[:n]
