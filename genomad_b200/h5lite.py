"""
h5lite -- a minimal, dependency-free reader for the subset of HDF5 that Keras'
legacy ``.h5`` weight files use (h5py is not a dependency of this package).

Why it exists: the reference loads ``genomad/data/nn_classifier.h5`` through
Keras' legacy-H5 loader (reference ``genomad/modules/nn_classification.py:309-310``,
``genomad/_paths.py:20-22``).  That file is HDF5 superblock version 0 with
old-style groups (B-tree ``TREE`` nodes + ``SNOD`` symbol nodes + a local
``HEAP``), version-1 object headers, and contiguous, unfiltered, little-endian
datasets.  This module parses exactly that subset and fails loudly on anything
else (chunked/compressed layouts, new-style groups, ...).

Public API
----------
``H5File(path)``                 parse the whole tree eagerly (files are small)
``H5File.datasets``              ``{"/model/conv1d/kernel:0": np.ndarray, ...}``
``H5File.attrs[path]``           ``{attr_name: value}`` for groups and datasets
``H5File.offsets[path]``         byte offset of each dataset's raw data
"""
from __future__ import annotations

import struct
from pathlib import Path
from typing import Dict, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class H5File:
    def __init__(self, path):
        self.path = Path(path)
        self.buf = self.path.read_bytes()
        self.datasets: Dict[str, np.ndarray] = {}
        self.offsets: Dict[str, int] = {}
        self.attrs: Dict[str, Dict[str, object]] = {}
        self._parse_superblock()

    # ------------------------------------------------------------------ low level
    def _u(self, fmt: str, off: int):
        return struct.unpack_from("<" + fmt, self.buf, off)

    def _parse_superblock(self) -> None:
        b = self.buf
        if b[:8] != _SIG:
            raise H5FormatError("not an HDF5 file")
        version = b[8]
        if version != 0:
            raise H5FormatError(f"superblock version {version} unsupported (need 0)")
        size_off, size_len = b[13], b[14]
        if (size_off, size_len) != (8, 8):
            raise H5FormatError("only 8-byte offsets/lengths supported")
        base, _free, eof, _drv = self._u("QQQQ", 24)
        if base != 0:
            raise H5FormatError("non-zero base address unsupported")
        if eof != len(b):
            raise H5FormatError(f"EOF address {eof} != file size {len(b)}")
        # root group symbol-table entry at byte 56
        _name_off, ohdr, cache_type, _ = self._u("QQII", 56)
        self._walk_object("", ohdr)

    # ------------------------------------------------------------------ objects
    def _messages(self, addr: int):
        """Yield (type, flags, payload_offset, payload_size) for a v1 object header."""
        version, _r, nmsg, _refc, hsize = self._u("BBHII", addr)
        if version != 1:
            raise H5FormatError(f"object header version {version} unsupported at {addr}")
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            off, size = blocks.pop(0)
            end = off + size
            while off + 8 <= end and seen < nmsg:
                mtype, msize, mflags = self._u("HHB", off)
                poff = off + 8
                seen += 1
                if mtype == 0x10:  # continuation
                    c_off, c_len = self._u("QQ", poff)
                    blocks.append((c_off, c_len))
                else:
                    yield mtype, mflags, poff, msize
                off = poff + msize

    def _walk_object(self, path: str, addr: int) -> None:
        space = dtype = layout = None
        symtab = None
        attrs: Dict[str, object] = {}
        for mtype, _flags, off, size in self._messages(addr):
            if mtype == 0x01:
                space = self._parse_dataspace(off)
            elif mtype == 0x03:
                dtype = self._parse_datatype(off)[0]
            elif mtype == 0x08:
                layout = self._parse_layout(off)
            elif mtype == 0x0B:
                raise H5FormatError(f"{path}: filter pipeline (compressed data) unsupported")
            elif mtype == 0x0C:
                name, val = self._parse_attribute(off)
                attrs[name] = val
            elif mtype == 0x11:
                symtab = self._u("QQ", off)
        self.attrs[path or "/"] = attrs
        if symtab is not None:
            btree, heap = symtab
            heap_data = self._heap_data_addr(heap)
            for name, child in self._iter_group(btree, heap_data):
                self._walk_object(f"{path}/{name}", child)
        elif layout is not None:
            if space is None or dtype is None:
                raise H5FormatError(f"{path}: dataset without dataspace/datatype")
            data_addr, data_size = layout
            count = int(np.prod(space)) if len(space) else 1
            nbytes = count * dtype.itemsize
            if data_addr == _UNDEF:
                arr = np.zeros(space, dtype=dtype)
            else:
                if data_size is not None and data_size != nbytes:
                    raise H5FormatError(f"{path}: layout size {data_size} != {nbytes}")
                arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=data_addr)
                arr = arr.reshape(space)
            self.datasets[path] = arr
            self.offsets[path] = data_addr

    # ------------------------------------------------------------------ groups
    def _heap_data_addr(self, addr: int) -> int:
        if self.buf[addr:addr + 4] != b"HEAP":
            raise H5FormatError(f"bad local heap signature at {addr}")
        _size, _free, data = self._u("QQQ", addr + 8)
        return data

    def _cstr(self, off: int) -> str:
        end = self.buf.index(b"\x00", off)
        return self.buf[off:end].decode("utf-8")

    def _iter_group(self, btree: int, heap_data: int):
        if self.buf[btree:btree + 4] != b"TREE":
            raise H5FormatError(f"bad B-tree signature at {btree}")
        ntype, level, used = self._u("BBH", btree + 4)
        if ntype != 0:
            raise H5FormatError("non-group B-tree where a group was expected")
        # keys and children interleave after the 24-byte header: k0 c0 k1 c1 ... kN
        base = btree + 24
        for i in range(used):
            child = self._u("Q", base + 8 + i * 16)[0]
            if level > 0:
                yield from self._iter_group(child, heap_data)
            else:
                if self.buf[child:child + 4] != b"SNOD":
                    raise H5FormatError(f"bad symbol node signature at {child}")
                nsym = self._u("H", child + 6)[0]
                for s in range(nsym):
                    e = child + 8 + s * 40
                    name_off, ohdr = self._u("QQ", e)
                    yield self._cstr(heap_data + name_off), ohdr

    def _global_heap_object(self, addr: int, idx: int) -> bytes:
        if self.buf[addr:addr + 4] != b"GCOL":
            raise H5FormatError(f"bad global heap signature at {addr}")
        size = self._u("Q", addr + 8)[0]
        off, end = addr + 16, addr + size
        while off + 16 <= end:
            oidx, _ref, _r, osize = self._u("HHIQ", off)
            if oidx == 0:
                break
            if oidx == idx:
                return self.buf[off + 16: off + 16 + osize]
            off += 16 + _pad8(osize)
        raise H5FormatError(f"global heap object {idx} not found at {addr}")

    # ------------------------------------------------------------------ messages
    def _parse_dataspace(self, off: int) -> Tuple[int, ...]:
        version, rank, flags = self._u("BBB", off)
        if version == 1:
            dims_off = off + 8
        elif version == 2:
            dims_off = off + 4
        else:
            raise H5FormatError(f"dataspace version {version} unsupported")
        return tuple(self._u(f"{rank}Q", dims_off)) if rank else ()

    def _parse_datatype(self, off: int):
        """Return (numpy dtype, total message bytes consumed)."""
        cv, b0, b1, _b2, size = self._u("BBBBI", off)
        cls, _ver = cv & 0x0F, cv >> 4
        if b0 & 1 and cls in (0, 1):
            raise H5FormatError("big-endian data unsupported")
        if cls == 0:  # fixed point: 4 bytes of properties
            signed = bool(b0 & 0x08)
            return np.dtype(f"<{'i' if signed else 'u'}{size}"), 8 + 4
        if cls == 1:  # floating point: 12 bytes of properties
            return np.dtype(f"<f{size}"), 8 + 12
        if cls == 3:  # fixed-length string
            return np.dtype(f"S{size}"), 8
        if cls == 9 and (b0 & 0x0F) == 1:  # variable-length string -> global heap refs
            return np.dtype([("len", "<u4"), ("addr", "<u8"), ("idx", "<u4")]), 8
        raise H5FormatError(f"datatype class {cls} unsupported")

    def _parse_layout(self, off: int):
        version = self.buf[off]
        if version == 3:
            lclass = self.buf[off + 1]
            if lclass != 1:
                raise H5FormatError(f"layout class {lclass} unsupported (need contiguous)")
            addr, size = self._u("QQ", off + 2)
            return addr, size
        if version in (1, 2):
            rank, lclass = self.buf[off + 1], self.buf[off + 2]
            if lclass != 1:
                raise H5FormatError(f"layout class {lclass} unsupported (need contiguous)")
            addr = self._u("Q", off + 8)[0]
            return addr, None
        raise H5FormatError(f"layout version {version} unsupported")

    def _parse_attribute(self, off: int):
        version = self.buf[off]
        if version == 1:
            name_sz, dt_sz, sp_sz = self._u("HHH", off + 2)
            p = off + 8
            name = self.buf[p:p + name_sz].split(b"\x00")[0].decode()
            p += _pad8(name_sz)
            dt_off = p
            p += _pad8(dt_sz)
            sp_off = p
            p += _pad8(sp_sz)
        elif version in (2, 3):
            name_sz, dt_sz, sp_sz = self._u("HHH", off + 2)
            p = off + 8 + (1 if version == 3 else 0)
            name = self.buf[p:p + name_sz].split(b"\x00")[0].decode()
            p += name_sz
            dt_off = p
            p += dt_sz
            sp_off = p
            p += sp_sz
        else:
            raise H5FormatError(f"attribute version {version} unsupported")
        try:
            dtype = self._parse_datatype(dt_off)[0]
        except H5FormatError:
            return name, None  # e.g. variable-length strings: not needed here
        shape = self._parse_dataspace(sp_off)
        count = int(np.prod(shape)) if len(shape) else 1
        arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=p).reshape(shape)
        if dtype.names:  # variable-length strings
            vals = [self._global_heap_object(int(r["addr"]), int(r["idx"]))[: int(r["len"])].decode()
                    for r in arr.reshape(-1)]
            return name, (vals if len(shape) else vals[0])
        if dtype.kind == "S":
            vals = [x.decode() for x in arr.reshape(-1)]
            return name, (vals if len(shape) else vals[0])
        return name, (arr.copy() if len(shape) else arr.reshape(-1)[0])
