"""
Counter-based synthetic 6 kb windows (BASELINE config 2: "generated on device from a counter-based generator keyed by
(seed, window index): uniform ACGT, plus a 1 % sub-stream with N runs / IUPAC for encoder coverage").

Window ``w`` of stream ``seed`` is a pure function of (seed, w): byte i is "ACGT"[h & 3] with
h = mix32(mix32(w ^ K(seed)) + i) (the lowbias32 integer hash); one window in 100 (mix32(w ^ K2) % 100 == 0) also carries a
300-byte run of 'N' and one IUPAC 'R'.  The same few integer expressions run on NumPy arrays (CPU oracle, fixtures) and on
torch CUDA tensors (bench.py, GPU tests), so any subsample of the 1 M-window stream can be regenerated anywhere
without storing it.  Nothing here is part of the classification path.
"""
from __future__ import annotations

import numpy as np

WINDOW = 6000
_M32 = 0xFFFFFFFF


def _mix32(x):
    """lowbias32 on non-negative int64 values holding 32-bit words (works for numpy arrays and torch tensors alike)."""
    x = x ^ (x >> 16)
    x = (x * 0x7FEB352D) & _M32
    x = x ^ (x >> 15)
    x = (x * 0x846CA68B) & _M32          # the int64 product may wrap; its low 32 bits are still exact
    x = x ^ (x >> 16)
    return x


def _keys(seed: int):
    k1 = (int(seed) * 0x9E3779B1 + 0x7F4A7C15) & _M32
    return k1, (k1 ^ 0x5BD1E995) & _M32, (k1 ^ 0x2545F491) & _M32


def windows_numpy(index, seed: int = 1) -> np.ndarray:
    """uint8 [len(index), 6000] for the given global window indices."""
    idx = np.asarray(index, dtype=np.int64).reshape(-1)
    k1, k2, k3 = _keys(seed)
    with np.errstate(over="ignore"):
        base = _mix32((idx & _M32) ^ k1)
        h = _mix32((base[:, None] + np.arange(WINDOW, dtype=np.int64)[None, :]) & _M32)
        a = np.frombuffer(b"ACGT", np.uint8)[h & 3].copy()
        dirty = _mix32((idx & _M32) ^ k2) % 100 == 0
        start = _mix32((idx & _M32) ^ k3) % 5500
    for r in np.nonzero(dirty)[0]:
        s = int(start[r])
        a[r, s:s + 300] = ord("N")
        a[r, (s * 7) % WINDOW] = ord("R")
    return a


def windows_torch(first: int, count: int, seed: int, device):
    """uint8 cuda tensor [count, 6000] holding windows first .. first+count-1 of the stream (generated on the device)."""
    import torch
    idx = torch.arange(first, first + count, dtype=torch.int64, device=device)
    k1, k2, k3 = _keys(seed)
    base = _mix32((idx & _M32) ^ k1)
    h = _mix32((base[:, None] + torch.arange(WINDOW, dtype=torch.int64, device=device)[None, :]) & _M32)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    a = lut[h & 3]
    dirty = torch.nonzero(_mix32((idx & _M32) ^ k2) % 100 == 0).flatten().tolist()
    start = (_mix32((idx & _M32) ^ k3) % 5500).tolist()
    for r in dirty:
        s = int(start[r])
        a[r, s:s + 300] = ord("N")
        a[r, (s * 7) % WINDOW] = ord("R")
    return a


def subsample_indices(n: int, total: int = 1_000_000, seed: int = 1) -> np.ndarray:
    """A fixed, seeded, sorted subsample of the window stream that always contains >= n // 64 "dirty" windows."""
    rng = np.random.default_rng(seed)
    k2 = _keys(seed)[1]
    with np.errstate(over="ignore"):
        all_dirty = np.nonzero(_mix32(np.arange(total, dtype=np.int64) ^ k2) % 100 == 0)[0]
    forced = rng.choice(all_dirty, size=max(1, n // 64), replace=False)
    rest = rng.choice(total, size=n, replace=False)
    merged = list(dict.fromkeys(list(forced) + list(rest)))[:n]
    return np.sort(np.asarray(merged, dtype=np.int64))
