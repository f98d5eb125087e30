"""Seeded synthetic single-cell Hi-C data of the shapes BASELINE.json names.

There is no network for real data, so the benchmark and the large parity cases use cells drawn
from a small generative model fitted to the reference's 20-cell fixture: mm9 chromosome layout of
``test_matrix.scool`` (2744 bins at 1 Mb), log-normal non-zeros per cell, 76 % intra-chromosomal
pixels with a power-law distance decay, a nearly complete main diagonal, and a per-cluster shared
"backbone" so that pairwise Jaccard similarities (hence MinHash collision counts) land where the
fixture's do. Every cell has its own PCG64 stream ``SeedSequence([20260319, config_id, cell_id])``,
so any rank can generate exactly its shard.
"""
from __future__ import annotations

import multiprocessing as mp
import os
from dataclasses import dataclass

import numpy as np
from scipy.sparse import csr_matrix

from .loader import ChromTable

# chromosome layout of the reference fixture (mm9, file order of test_matrix.scool)
MM9 = [('chr10', 129993255), ('chr11', 121843856), ('chr12', 121257530), ('chr13', 120284312),
       ('chr13_random', 400311), ('chr14', 125194864), ('chr15', 103494974), ('chr16', 98319150),
       ('chr16_random', 3994), ('chr17', 95272651), ('chr17_random', 628739), ('chr18', 90772031),
       ('chr19', 61342430), ('chr1', 197195432), ('chr1_random', 1231697), ('chr2', 181748087),
       ('chr3', 159599783), ('chr3_random', 41899), ('chr4', 155630120), ('chr4_random', 160594),
       ('chr5', 152537259), ('chr5_random', 357350), ('chr6', 149517037), ('chr7', 152524553),
       ('chr7_random', 362490), ('chr8', 131738871), ('chr8_random', 849593), ('chr9', 124076172),
       ('chr9_random', 449403), ('chrM', 16299), ('chrUn_random', 5900358), ('chrX', 166650296),
       ('chrX_random', 1785075), ('chrY', 15902555), ('chrY_random', 58682461)]

BASE_SEED = 20260319


@dataclass(frozen=True)
class SynthConfig:
    config_id: int
    n_cells: int
    binsize: int = 1_000_000
    median_nnz: int = 15_000
    sigma: float = 0.8
    min_nnz: int = 1_500
    max_nnz: int = 60_000
    intra_fraction: float = 0.76      # 1.0 = intra-chromosomal only
    n_clusters: int = 3
    chromosomes: tuple = ()           # restrict bin1 to these chromosomes (config 5)

    @property
    def chroms(self) -> ChromTable:
        return ChromTable.from_sizes([n for n, _ in MM9], [s for _, s in MM9], self.binsize)


# the configurations of BASELINE.json / BASELINE.md section 3
CONFIGS = {
    "nagano": SynthConfig(2, 3882, median_nnz=12_000, intra_fraction=1.0, n_clusters=7),
    "s10": SynthConfig(3, 10_000),
    "s50": SynthConfig(4, 50_000, binsize=500_000, median_nnz=22_000),
    "sweep": SynthConfig(5, 10_000, chromosomes=("chr1", "chr2")),
}


def small_config(n_cells: int, config_id: int = 99, **kw) -> SynthConfig:
    """Reduced-size cells for parity tests (same model, fewer non-zeros)."""
    base = dict(median_nnz=3_000, min_nnz=200, max_nnz=12_000)
    base.update(kw)
    return SynthConfig(config_id, n_cells, **base)


def _rng(cfg: SynthConfig, cell_id: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(np.random.SeedSequence([BASE_SEED, cfg.config_id, int(cell_id)])))


_DMAX = 400


def _distance_cdf(alpha: float) -> np.ndarray:
    p = (np.arange(_DMAX) + 1.0) ** (-alpha)
    return np.cumsum(p / p.sum())


def _draw_pixels(rng, cfg, chroms, n, alpha):
    """n (bin1, bin2) pixels: intra with distance ~ (d+1)^-alpha, the rest uniform inter."""
    nb = np.diff(chroms.offsets)
    n_intra = int(rng.binomial(n, cfg.intra_fraction)) if cfg.intra_fraction < 1.0 else n
    ch = rng.choice(len(nb), size=n_intra, p=nb / nb.sum())
    b1 = chroms.offsets[ch] + (rng.random(n_intra) * nb[ch]).astype(np.int64)
    d = np.searchsorted(_distance_cdf(alpha), rng.random(n_intra))
    b2 = np.minimum(b1 + d, chroms.offsets[ch + 1] - 1)
    n_inter = n - n_intra
    if n_inter > 0:
        x = rng.integers(0, chroms.nbins, size=(n_inter, 2))
        b1 = np.concatenate([b1, x.min(axis=1)])
        b2 = np.concatenate([b2, x.max(axis=1)])
    return b1, b2


def _backbone(cfg: SynthConfig, chroms: ChromTable, z: int) -> np.ndarray:
    rng = _rng(cfg, (1 << 31) + z)
    alpha = 0.9 + 0.4 * (z / max(1, cfg.n_clusters - 1))
    b1, b2 = _draw_pixels(rng, cfg, chroms, int(0.35 * cfg.median_nnz), alpha)
    return np.unique(b1 * np.int64(chroms.nbins) + b2)


def generate_cell(cfg: SynthConfig, cell_id: int, chroms=None, backbones=None):
    """One cell: sorted unique feature ids (int64) and integer counts (float32)."""
    chroms = chroms or cfg.chroms
    rng = _rng(cfg, cell_id)
    z = int(rng.integers(0, cfg.n_clusters))
    alpha = 0.9 + 0.4 * (z / max(1, cfg.n_clusters - 1))
    target = int(np.clip(rng.lognormal(np.log(cfg.median_nnz), cfg.sigma), cfg.min_nnz, cfg.max_nnz))
    nbins = chroms.nbins
    diag = np.nonzero(rng.random(nbins) < 0.94)[0].astype(np.int64)
    bb = backbones[z] if backbones is not None else _backbone(cfg, chroms, z)
    bb = bb[rng.random(len(bb)) < 0.5]
    feats = [diag * nbins + diag, bb]
    rest = target - len(diag) - len(bb)
    if rest > 0:
        b1, b2 = _draw_pixels(rng, cfg, chroms, rest, alpha)
        feats.append(b1 * np.int64(nbins) + b2)
    f = np.unique(np.concatenate(feats))
    if len(f) > target:   # small targets: thin uniformly instead of exceeding the drawn size
        f = np.sort(rng.choice(f, size=target, replace=False))
    if cfg.chromosomes:   # --chromosomes keeps pixels whose bin1 lies in the listed chromosomes
        idx = {n: i for i, n in enumerate(chroms.names)}
        c1 = chroms.chrom_of_bin(f // nbins)
        f = f[np.isin(c1, [idx[c] for c in cfg.chromosomes])]
    cnt = np.minimum(rng.geometric(0.3, size=len(f)), 600).astype(np.float32)
    return f, cnt, z


def _gen_range(args):
    cfg, lo, hi = args
    chroms = cfg.chroms
    backbones = [_backbone(cfg, chroms, z) for z in range(cfg.n_clusters)]
    fs, cs, zs = [], [], []
    for cid in range(lo, hi):
        f, c, z = generate_cell(cfg, cid, chroms, backbones)
        fs.append(f.astype(np.uint32) if chroms.nbins ** 2 <= 0xFFFFFFFF else f)
        cs.append(c)
        zs.append(z)
    return fs, cs, zs


def generate_csr(cfg: SynthConfig, lo: int = 0, hi: int | None = None, workers: int | None = None):
    """CSR pieces of cells [lo, hi): ``indptr`` int64, ``indices`` uint32 (feature ids), ``data`` float32,
    plus the latent cluster labels. Deterministic whatever the worker count."""
    hi = cfg.n_cells if hi is None else hi
    n = hi - lo
    workers = workers or min(32, os.cpu_count() or 1)
    step = max(1, (n + workers * 4 - 1) // (workers * 4))
    jobs = [(cfg, a, min(hi, a + step)) for a in range(lo, hi, step)]
    if workers > 1 and n >= 64:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_gen_range, jobs)
    else:
        parts = [_gen_range(j) for j in jobs]
    fs = [f for p in parts for f in p[0]]
    cs = [c for p in parts for c in p[1]]
    zs = np.array([z for p in parts for z in p[2]], dtype=np.int32)
    indptr = np.concatenate([[0], np.cumsum([len(f) for f in fs])]).astype(np.int64)
    indices = np.concatenate(fs) if fs else np.zeros(0, dtype=np.uint32)
    data = np.concatenate(cs) if cs else np.zeros(0, dtype=np.float32)
    return indptr, indices, data, zs


def to_scipy(cfg: SynthConfig, indptr, indices, data) -> csr_matrix:
    nb = cfg.chroms.nbins
    idx = indices.astype(np.int32) if nb * nb < 2 ** 31 else indices.astype(np.int64)
    return csr_matrix((data.astype(np.float64), idx, indptr), shape=(len(indptr) - 1, nb * nb))
