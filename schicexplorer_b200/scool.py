"""Cooler-free reader for ``.scool`` files.

``.scool`` is HDF5 in the cooler multi-cell layout; neither h5py, cooler nor libhdf5 is installed here,
and the only things the clustering path needs are, per cell, the three pixel columns
(``/cells/<name>/pixels/{bin1_id,bin2_id,count}``), the shared bin table (``/bins``) and the chromosome
table (``/chroms``) -- what ``load_matrix`` pulls through cooler in the reference
(``/root/reference/schicexplorer/utilities.py:54-76``, cell order from ``cell_name_list`` :31-35).
This module reads exactly the HDF5 subset such files use (format spec, "classic" layout written by
h5py/libhdf5 1.8+ defaults):

* superblock version 0/1, 8-byte offsets and lengths;
* version-1 object headers (with continuation blocks), symbol-table groups (v1 B-tree + local heap);
* datasets: contiguous or chunked (v1 chunk B-tree), fixed-point / fixed-length-string / enum datatypes,
  filters deflate (1) and shuffle (2).

Anything else raises ``ScoolFormatError`` -- no silent guesses.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

from .loader import CellPixels, ChromTable

UNDEF = 0xFFFFFFFFFFFFFFFF


class ScoolFormatError(RuntimeError):
    pass


class _Dataset:
    def __init__(self, shape, dtype, layout, filters, enum=None):
        self.shape, self.dtype, self.layout, self.filters, self.enum = shape, dtype, layout, filters, enum


class HDF5File:
    """Minimal read-only HDF5 (see module docstring for the supported subset)."""

    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise ScoolFormatError(f"{path}: not an HDF5 file")
        ver = self.buf[8]
        if ver not in (0, 1):
            raise ScoolFormatError(f"superblock version {ver} is not supported (only 0/1)")
        so, sl = self.buf[13], self.buf[14]
        if (so, sl) != (8, 8):
            raise ScoolFormatError("only 8-byte offsets/lengths are supported")
        off = 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", self.buf, off)[0]
        root_entry = off + 32
        self.root = self._read_symbol_entry(root_entry)[1]

    # ---- low-level structures -------------------------------------------------------------------
    def _read_symbol_entry(self, pos):
        name_off, ohdr, cache = struct.unpack_from("<QQI", self.buf, pos)
        return name_off, ohdr + self.base, cache

    def _messages(self, addr):
        """(type, flags, payload) of a version-1 object header, following continuation blocks."""
        b = self.buf
        if b[addr] != 1:
            raise ScoolFormatError(f"object header version {b[addr]} at {addr} is not supported (only 1)")
        nmsg, = struct.unpack_from("<H", b, addr + 2)
        size, = struct.unpack_from("<I", b, addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, pos)
                body = b[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:      # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr + self.base, clen))
                out.append((mtype, flags, body))
        return out

    def _heap_data(self, addr):
        if self.buf[addr:addr + 4] != b"HEAP":
            raise ScoolFormatError("bad local heap signature")
        data_addr, = struct.unpack_from("<Q", self.buf, addr + 24)
        return data_addr + self.base

    def _cstr(self, pos):
        end = self.buf.index(b"\x00", pos)
        return self.buf[pos:end].decode("ascii")

    def _group_entries(self, btree, heap):
        """name -> object header address of a symbol-table group, in B-tree (name) order."""
        heap_data = self._heap_data(heap)
        out = []

        def walk(addr):
            b = self.buf
            if b[addr:addr + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
                if ntype != 0:
                    raise ScoolFormatError("expected a group B-tree node")
                pos = addr + 24            # signature, type, level, used, left, right
                for i in range(used):
                    child, = struct.unpack_from("<Q", b, pos + 8 + i * 16)
                    walk(child + self.base)
            elif b[addr:addr + 4] == b"SNOD":
                n, = struct.unpack_from("<H", b, addr + 6)
                for i in range(n):
                    name_off, ohdr, _ = self._read_symbol_entry(addr + 8 + i * 40)
                    out.append((self._cstr(heap_data + name_off), ohdr))
            else:
                raise ScoolFormatError(f"unexpected node signature at {addr}")
        walk(btree)
        return out

    def listdir(self, addr):
        for mtype, _, body in self._messages(addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                return self._group_entries(btree + self.base, heap + self.base)
        raise ScoolFormatError("object is not an (old-style) group")

    def resolve(self, path: str, start=None):
        addr = self.root if start is None else start
        for part in [p for p in path.split("/") if p]:
            entries = dict(self.listdir(addr))
            if part not in entries:
                raise KeyError(path)
            addr = entries[part]
        return addr

    # ---- datasets ---------------------------------------------------------------------------------
    @staticmethod
    def _parse_datatype(body):
        cls = body[0] & 0x0F
        ver = body[0] >> 4
        bits0 = body[1]
        size, = struct.unpack_from("<I", body, 4)
        if cls == 0:                                   # fixed point
            order = ">" if (bits0 & 1) else "<"
            signed = bool(bits0 & 8)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), None, 8 + 4
        if cls == 1:                                   # floating point
            order = ">" if (bits0 & 1) else "<"
            return np.dtype(f"{order}f{size}"), None, 8 + 12
        if cls == 3:                                   # fixed-length string
            return np.dtype(f"S{size}"), None, 8
        if cls == 8:                                   # enum over an integer base type
            nmembers = body[1] | (body[2] << 8)
            base, _, blen = HDF5File._parse_datatype(body[8:])
            pos = 8 + blen
            names = []
            for _ in range(nmembers):
                end = body.index(b"\x00", pos)
                names.append(body[pos:end].decode("ascii"))
                ln = end - pos + 1
                pos += ln if ver >= 3 else (ln + 7) // 8 * 8
            values = np.frombuffer(body, dtype=base, count=nmembers, offset=pos)
            return base, dict(zip(values.tolist(), names)), pos + nmembers * base.itemsize
        raise ScoolFormatError(f"datatype class {cls} is not supported")

    def dataset(self, addr) -> _Dataset:
        shape = dtype = layout = enum = None
        filters = []
        for mtype, _, body in self._messages(addr):
            if mtype == 0x0001:                        # dataspace
                ver, rank, flags = body[0], body[1], body[2]
                pos = 8 if ver == 1 else 4
                shape = struct.unpack_from(f"<{rank}Q", body, pos)
            elif mtype == 0x0003:
                dtype, enum, _ = self._parse_datatype(body)
            elif mtype == 0x0008:                      # layout
                ver, cls = body[0], body[1]
                if ver != 3:
                    raise ScoolFormatError(f"data layout version {ver} is not supported (only 3)")
                if cls == 1:
                    a, sz = struct.unpack_from("<QQ", body, 2)
                    layout = ("contiguous", a, sz)
                elif cls == 2:
                    rank = body[2]
                    a, = struct.unpack_from("<Q", body, 3)
                    dims = struct.unpack_from(f"<{rank}I", body, 11)
                    layout = ("chunked", a, dims)
                elif cls == 0:
                    sz, = struct.unpack_from("<H", body, 2)
                    layout = ("compact", bytes(body[4:4 + sz]))
            elif mtype == 0x000B:                      # filter pipeline
                ver, nf = body[0], body[1]
                pos = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid, nlen, _flags, ncd = struct.unpack_from("<HHHH", body, pos)
                    pos += 8
                    if ver == 1 or fid >= 256:
                        pos += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    cd = struct.unpack_from(f"<{ncd}I", body, pos)
                    pos += 4 * ncd
                    if ver == 1 and ncd % 2:
                        pos += 4
                    filters.append((fid, cd))
        if shape is None or dtype is None or layout is None:
            raise ScoolFormatError("object is not a dataset")
        return _Dataset(shape, dtype, layout, filters, enum)

    def _unfilter(self, raw, filters, mask, itemsize):
        for i, (fid, cd) in reversed(list(enumerate(filters))):
            if mask & (1 << i):
                continue
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                n = len(raw) // itemsize
                raw = np.ascontiguousarray(np.frombuffer(raw, dtype=np.uint8, count=n * itemsize).reshape(itemsize, n).T)
            else:
                raise ScoolFormatError(f"HDF5 filter {fid} is not supported (only deflate and shuffle)")
        return raw

    def read(self, addr) -> np.ndarray:
        ds = self.dataset(addr)
        n = int(np.prod(ds.shape)) if ds.shape else 1
        kind = ds.layout[0]
        if kind == "compact":
            return np.frombuffer(ds.layout[1], dtype=ds.dtype, count=n).reshape(ds.shape).copy()
        if kind == "contiguous":
            a = ds.layout[1]
            if a == UNDEF:
                return np.zeros(ds.shape, dtype=ds.dtype)
            return np.frombuffer(self.buf, dtype=ds.dtype, count=n, offset=a + self.base).reshape(ds.shape).copy()
        # chunked, rank 1 is all the cooler layout uses for the columns we read
        if len(ds.shape) != 1:
            raise ScoolFormatError("only 1-D chunked datasets are supported")
        out = np.zeros(ds.shape, dtype=ds.dtype)
        a, dims = ds.layout[1], ds.layout[2]
        chunk_elems = dims[0]
        if a == UNDEF:
            return out

        def walk(addr):
            b = self.buf
            if b[addr:addr + 4] != b"TREE":
                raise ScoolFormatError("bad chunk B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
            if ntype != 1:
                raise ScoolFormatError("expected a chunk B-tree node")
            key_size = 8 + 8 * (len(ds.shape) + 1)
            pos = addr + 24
            for i in range(used):
                csize, cmask = struct.unpack_from("<II", b, pos)
                offs = struct.unpack_from(f"<{len(ds.shape) + 1}Q", b, pos + 8)
                child, = struct.unpack_from("<Q", b, pos + key_size)
                pos += key_size + 8
                if level > 0:
                    walk(child + self.base)
                else:
                    raw = bytes(b[child + self.base: child + self.base + csize])
                    raw = self._unfilter(raw, ds.filters, cmask, ds.dtype.itemsize)
                    vals = np.frombuffer(raw, dtype=ds.dtype, count=chunk_elems)
                    s = offs[0]
                    e = min(s + chunk_elems, ds.shape[0])
                    out[s:e] = vals[:e - s]
        walk(a + self.base)
        return out


def _cell_groups(f: HDF5File):
    """(reported name, group address, address holding bins/chroms) per cell, in file order.

    New-style files keep the cells under ``/cells`` and share ``/bins`` and ``/chroms``
    (``cooler.fileops.list_scool_cells``); the old, non-standard layout stores one complete cooler
    per cell in the root group and is reported as ``/<name>`` (``cell_name_list``,
    ``/root/reference/schicexplorer/utilities.py:37-44``). Members of an HDF5 symbol-table group
    iterate in name order, which is the order cooler reports."""
    root = dict(f.listdir(f.root))
    if "cells" in root:
        return [("/cells/" + name, addr, f.root) for name, addr in f.listdir(root["cells"])]
    out = []
    for name, addr in f.listdir(f.root):
        try:
            members = dict(f.listdir(addr))
        except ScoolFormatError:
            continue
        if "pixels" in members and "bins" in members and "chroms" in members:
            out.append(("/" + name, addr, addr))
    if not out:
        raise ScoolFormatError("Wrong data format. Please use a scool file.")
    return out


def list_scool_cells(f) -> list:
    """Cell names in file order, as ``cell_name_list`` reports them (utilities.py:31-51)."""
    if not isinstance(f, HDF5File):
        f = HDF5File(f)
    return [name for name, _, _ in _cell_groups(f)]


def read_scool(path: str, cell_range=None, threads: int = 1):
    """Cells of a ``.scool`` file as :class:`CellPixels` + the :class:`ChromTable`.
    ``cell_range=(a, b)`` restricts the pixel tables read to cells [a, b) (the --saveMemory chunks).
    ``threads`` > 1 reads the cells' pixel tables concurrently (the reference fans the same work out over
    ``--threads`` processes, utilities.py:169-191; inflate and the numpy copies release the GIL)."""
    f = HDF5File(path)
    groups = _cell_groups(f)
    meta = groups[0][2]
    names = [s.decode("ascii") if isinstance(s, bytes) else str(s)
             for s in f.read(f.resolve("chroms/name", start=meta)).tolist()]
    lengths = f.read(f.resolve("chroms/length", start=meta)).astype(np.int64)
    start = f.read(f.resolve("bins/start", start=meta)).astype(np.int64)
    end = f.read(f.resolve("bins/end", start=meta)).astype(np.int64)
    binsize = int(end[0] - start[0]) if len(start) else 1
    chroms = ChromTable.from_sizes(names, lengths, binsize)
    if chroms.nbins != len(start):
        raise ScoolFormatError(f"bin table has {len(start)} bins, chromosome sizes imply {chroms.nbins}")
    if cell_range is not None:
        groups = groups[cell_range[0]:cell_range[1]]
    def one(group):
        name, addr, _ = group
        px = f.resolve("pixels", start=addr)
        return CellPixels(name,
                          f.read(f.resolve("bin1_id", start=px)).astype(np.int64),
                          f.read(f.resolve("bin2_id", start=px)).astype(np.int64),
                          f.read(f.resolve("count", start=px)).astype(np.float64))

    if threads and threads > 1 and len(groups) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(int(threads), len(groups))) as pool:
            cells = list(pool.map(one, groups))          # map keeps the file order
    else:
        cells = [one(g) for g in groups]
    return cells, chroms
