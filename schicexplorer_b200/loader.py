"""Host-side input contract of the hot path: per-cell pixel tables -> one CSR (cells x nbins^2).

Restates the semantics of the reference loader (``/root/reference/schicexplorer/utilities.py``:
``load_matrix`` :54-76, ``open_and_store_matrix`` :79-138, ``create_csr_matrix_all_cells`` :141-234)
without cooler/HDF5 (neither is installed here): the per-cell pixel tables come from the
``sparse-matrix-files`` text dump that ``scHicConvertFormat`` writes (scHicConvertFormat.py:79-99),
from ``.npz`` interchange files, or from the synthetic generator. Loading stays on the host; only
the resulting CSR crosses the C ABI.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
from scipy.sparse import csr_matrix


@dataclass
class CellPixels:
    """Upper-triangular pixel table of one cell (cooler ``pixels`` columns)."""
    name: str
    bin1: np.ndarray   # int64
    bin2: np.ndarray   # int64
    count: np.ndarray  # float64


@dataclass
class ChromTable:
    names: list
    sizes: np.ndarray      # bp
    binsize: int
    offsets: np.ndarray    # first bin of each chromosome, len = n_chrom + 1

    @property
    def nbins(self) -> int:
        return int(self.offsets[-1])

    @classmethod
    def from_sizes(cls, names, sizes, binsize):
        sizes = np.asarray(sizes, dtype=np.int64)
        nb = (sizes + binsize - 1) // binsize
        return cls(list(names), sizes, int(binsize), np.concatenate([[0], np.cumsum(nb)]).astype(np.int64))

    def chrom_of_bin(self, bins: np.ndarray) -> np.ndarray:
        return np.searchsorted(self.offsets, bins, side="right") - 1


def read_chromosome_sizes(path: str, binsize: int) -> ChromTable:
    names, sizes = [], []
    with open(path) as fh:
        for line in fh:
            if line.strip():
                a, b = line.split()[:2]
                names.append(a)
                sizes.append(int(b))
    return ChromTable.from_sizes(names, sizes, binsize)


def read_sparse_matrix_files(folder: str, binsize: int = 1_000_000):
    """Read a ``sparse-matrix-files`` dump: ``cellNameFile.txt`` gives the row order (= the order of
    ``cell_name_list``, utilities.py:31-35), ``cells/<name>.txt`` holds ``bin1 bin2 count``."""
    chroms = read_chromosome_sizes(os.path.join(folder, "chromosomeSize.txt"), binsize)
    cells = []
    with open(os.path.join(folder, "cellNameFile.txt")) as fh:
        names = [l.strip() for l in fh if l.strip()]
    for n in names:
        base = os.path.basename(n)
        tbl = np.loadtxt(os.path.join(folder, "cells", base + ".txt"), dtype=np.int64, ndmin=2)
        if tbl.size == 0:
            tbl = np.zeros((0, 3), dtype=np.int64)
        cells.append(CellPixels("/cells/" + base, tbl[:, 0].copy(), tbl[:, 1].copy(), tbl[:, 2].astype(np.float64)))
    return cells, chroms


def select_pixels(cell: CellPixels, chroms: ChromTable, chromosomes=None, intra_only=False, distance_bins=None):
    """Pixel filter of ``load_matrix`` / ``open_and_store_matrix``:
    * ``chromosomes``: keep pixels whose *bin1* lies in one of them, concatenated in the order
      given (``pixels().fetch(chrom)``, utilities.py:60-64); default = all, file order.
    * ``intra_only`` (-ic): additionally bin2 in the same chromosome (utilities.py:68-70).
    * ``distance_bins`` (-d, already divided by the bin size, utilities.py:167-168): keep
      ``|bin1-bin2| < d`` and only if ``d > 5`` -- a scalar test, so d <= 5 drops everything
      (utilities.py:106-113)."""
    if chromosomes is None and not intra_only and distance_bins is None:
        return cell.bin1, cell.bin2, cell.count          # nothing to filter: no per-pixel chromosome lookup
    c1 = chroms.chrom_of_bin(cell.bin1)
    if chromosomes is None:
        parts = [np.arange(len(cell.bin1))]
    else:
        idx = {n: i for i, n in enumerate(chroms.names)}
        parts = [np.nonzero(c1 == idx[ch])[0] for ch in chromosomes]
    sel = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)
    if intra_only:
        c2 = chroms.chrom_of_bin(cell.bin2[sel])
        sel = sel[c2 == c1[sel]]
    b1, b2, cnt = cell.bin1[sel], cell.bin2[sel], cell.count[sel]
    if distance_bins is not None:
        mask = (np.abs(b1 - b2) < distance_bins) & (distance_bins > 5)
        b1, b2, cnt = b1[mask], b2[mask], cnt[mask]
    return b1, b2, cnt


def cells_to_csr(cells, chroms: ChromTable, chromosomes=None, intra_only=False, distance=None) -> tuple:
    """``create_csr_matrix_all_cells`` (utilities.py:141-234): rows in cell order, feature id
    ``bin1 * nbins + bin2`` with the FULL-genome bin count (utilities.py:119-120, :131), values =
    raw counts as float64. Deviation (documented in SURVEY appendix B.5): an empty cell keeps its
    own all-zero row instead of shifting the later cells of its worker up by one.
    Returns ``(csr, names)``."""
    nbins = chroms.nbins
    dbins = None if distance is None else distance // chroms.binsize
    counts = np.zeros(len(cells) + 1, dtype=np.int64)
    feats, vals = [], []
    for i, cell in enumerate(cells):
        b1, b2, cnt = select_pixels(cell, chroms, chromosomes, intra_only, dbins)
        f = b1 * np.int64(nbins) + b2
        # what the COO -> CSR constructor of utilities.py:131 does per row (duplicates summed, columns sorted), done
        # row by row: pixel tables come sorted and duplicate free from a cooler, so this is normally just a check
        if len(f) > 1 and not np.all(f[1:] > f[:-1]):
            order = np.argsort(f, kind="stable")
            f, cnt = f[order], np.asarray(cnt, dtype=np.float64)[order]
            first = np.concatenate([[True], f[1:] != f[:-1]])
            cnt = np.add.reduceat(cnt, np.nonzero(first)[0])
            f = f[first]
        feats.append(f)
        vals.append(np.asarray(cnt, dtype=np.float64))
        counts[i + 1] = len(f)
    indptr = np.cumsum(counts)
    f = np.concatenate(feats) if feats else np.zeros(0, dtype=np.int64)
    v = np.concatenate(vals) if vals else np.zeros(0)
    shape = (len(cells), nbins * nbins)
    # index dtype as scipy picks it for this shape (int32 while nbins^2 < 2^31: SURVEY 8(c))
    idx_dtype = np.int32 if max(shape) < 2 ** 31 and indptr[-1] < 2 ** 31 else np.int64
    m = csr_matrix((v, f.astype(idx_dtype, copy=False), indptr.astype(idx_dtype)), shape=shape, dtype=np.float64)
    m.has_sorted_indices = True
    m.has_canonical_format = True
    return m, [c.name for c in cells]


def save_cells_npz(path: str, cells, chroms: ChromTable):
    offs = np.concatenate([[0], np.cumsum([len(c.bin1) for c in cells])]).astype(np.int64)
    np.savez_compressed(
        path, names=np.array([c.name for c in cells]), offsets=offs,
        bin1=np.concatenate([c.bin1 for c in cells]).astype(np.int32),
        bin2=np.concatenate([c.bin2 for c in cells]).astype(np.int32),
        count=np.concatenate([c.count for c in cells]).astype(np.int32),
        chrom_names=np.array(chroms.names), chrom_sizes=chroms.sizes, binsize=np.int64(chroms.binsize))


def load_cells_npz(path: str):
    z = np.load(path, allow_pickle=False)
    chroms = ChromTable.from_sizes([str(s) for s in z["chrom_names"]], z["chrom_sizes"], int(z["binsize"]))
    offs = z["offsets"]
    cells = []
    for i, n in enumerate(z["names"]):
        a, b = offs[i], offs[i + 1]
        cells.append(CellPixels(str(n), z["bin1"][a:b].astype(np.int64), z["bin2"][a:b].astype(np.int64),
                                z["count"][a:b].astype(np.float64)))
    return cells, chroms
