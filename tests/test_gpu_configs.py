"""GPU suite (-m gpu): the BASELINE.json configurations.

Reduced-size instances of configs 2-5 (same generators, same per-cell sizes, fewer cells) against the oracle on
identical inputs, and the FULL config-3 size (10 000 cells, 1.43e8 non-zeros, H = 800, k = 100, exact rerank)
through size-independent properties plus a sampled check against the oracle / a direct sparse computation.
Bar as everywhere: bit-exact signatures, counts and indices; 1e-5 relative on reranked distances."""
import os

import numpy as np
import pytest

from conftest import REF_KWARGS, Fitted
from schicexplorer_b200 import synth
from test_gpu_parity import RTOL, assert_neighbours_equal_up_to_ties

pytestmark = pytest.mark.gpu


def cells(name, n, **over):
    cfg = synth.SynthConfig(**{**synth.CONFIGS[name].__dict__, "n_cells": n, **over})
    ip, ix, dt, _ = synth.generate_csr(cfg, 0, n, workers=min(16, os.cpu_count() or 1))
    return cfg, ip, ix, dt, synth.to_scipy(cfg, ip, ix, dt)


def test_config2_nagano_shape_dense_graph(cuda, oracle):
    """Config 2: Nagano-shaped cells (intra-chromosomal only, ~12 k non-zeros), -nh 800, k = n (the CLI default:
    the kNN graph is dense), fast distances. Graph identical to the oracle's."""
    _, _, _, _, X = cells("nagano", 320)
    n = X.shape[0]
    with Fitted(cuda, X, fast=True, n_neighbors=n) as g, Fitted(oracle, X, fast=True, n_neighbors=n) as o:
        assert np.array_equal(g.sig(), o.sig())
        gi, gc, gd = cuda.kneighbors_graph(g.h, n)
        oi, oc, od = oracle.kneighbors_graph(o.h, n)
        assert np.array_equal(gi, oi) and np.array_equal(gc, oc) and np.array_equal(gd, od)
        assert gi[-1] > n          # more than the self edges: clusters really collide


def test_config3_s10_shape_exact_rerank(cuda, oracle):
    """Config 3 at 360 cells: full-length rows, k = 100, exact euclidean rerank."""
    _, _, _, _, X = cells("s10", 360)
    with Fitted(cuda, X, fast=False, n_neighbors=100) as g, Fitted(oracle, X, fast=False, n_neighbors=100) as o:
        assert np.array_equal(g.sig(), o.sig())
        assert np.array_equal(cuda.collision_counts(g.h), oracle.collision_counts(o.h))
        a, b = cuda.kneighbors(g.h, 100), oracle.kneighbors(o.h, 100)
        assert_neighbours_equal_up_to_ties(*a, *b, RTOL)


def test_config4_s50_shape_500kb(cuda, oracle):
    """Config 4 at 300 cells: 500 kb bins (5 4xx bins, ids up to 3.0e7), ~22 k non-zeros per cell, fast, k = 100."""
    cfg, _, ix, _, X = cells("s50", 300)
    assert cfg.chroms.nbins > 5000 and int(ix.max()) > 2 ** 24
    with Fitted(cuda, X, fast=True, n_neighbors=100) as g, Fitted(oracle, X, fast=True, n_neighbors=100) as o:
        assert np.array_equal(g.sig(), o.sig())
        a, b = cuda.kneighbors(g.h, 100), oracle.kneighbors(o.h, 100)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_config5_hash_count_sweep_chromosome_subset(cuda, oracle):
    """Config 5: --chromosomes chr1 chr2, -nh 200 ... 2000. h_j depends only on j, so the signatures for a smaller
    hash count are a prefix of those for a larger one (the property the hyperopt sweep can exploit)."""
    _, _, _, _, X = cells("sweep", 200)
    sigs = {}
    for H in (200, 800, 2000):
        kw = dict(number_of_hash_functions=H, minimal_blocks_in_common=H // 8)
        with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
            sigs[H] = g.sig()
            assert np.array_equal(sigs[H], o.sig())
            a, b = cuda.kneighbors(g.h, 20), oracle.kneighbors(o.h, 20)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(sigs[2000][:, :200], sigs[200]) and np.array_equal(sigs[2000][:, :800], sigs[800])


def test_config3_full_size(cuda, oracle, monkeypatch):
    """The configuration the metric is quoted on, at full size: 10 000 cells, H = 800, k = 100, exact rerank.
    * K1: the shared-table path (chosen automatically at this size) and the per-cell filtered kernel give the same
      signature matrix; 48 sampled rows equal the oracle's signatures of those rows.
    * counts: packed (fp16-id) and plain K3 agree on a 256-row block against all 10 000 columns; the block is
      symmetric and its diagonal = number of admissible signature entries.
    * K4: windowed and hash-table kernels agree (indices identical, distances to 1e-6); every list starts with the
      query itself at distance 0 and is sorted; 32 sampled lists are recomputed directly from the sparse rows in
      float64 (1e-5 relative)."""
    cfg, ip, ix, dt, _ = cells("s10", 10_000)
    n, H, k = 10_000, 800, 100
    rng = np.random.default_rng(5)

    def build(k1_mode, k3_mode, k4_mode):
        monkeypatch.setenv("SCHIC_K1_MODE", k1_mode)
        monkeypatch.setenv("SCHIC_K3_MODE", k3_mode)
        monkeypatch.setenv("SCHIC_K4_MODE", k4_mode)
        h = cuda.create(n_neighbors=k, fast=False, **REF_KWARGS)
        cuda.fit_csr(h, ip, ix, dt, False)
        res = cuda.kneighbors(h, k)
        sig = cuda.signatures(h, H)
        cnt = cuda.collision_counts(h, 1000, 1256)
        cuda.destroy(h)
        return sig, res, cnt

    sig_a, (idx_a, dist_a, nv_a), cnt_a = build("auto", "packed", "window")
    sig_b, (idx_b, dist_b, nv_b), cnt_b = build("tile", "plain", "hash")
    assert np.array_equal(sig_a, sig_b) and np.array_equal(cnt_a, cnt_b)
    assert np.array_equal(nv_a, nv_b) and np.array_equal(idx_a, idx_b)
    np.testing.assert_allclose(dist_a, dist_b, rtol=1e-6)

    rows = np.sort(rng.choice(n, size=48, replace=False))
    sub_ip = np.concatenate([[0], np.cumsum(ip[rows + 1] - ip[rows])]).astype(np.int64)
    sub_ix = np.concatenate([ix[ip[r]:ip[r + 1]] for r in rows])
    ho = oracle.create(n_neighbors=5, fast=True, **REF_KWARGS)
    oracle.fit_csr(ho, sub_ip, sub_ix, None, False)
    assert np.array_equal(oracle.signatures(ho, H), sig_a[rows])
    oracle.destroy(ho)

    blk = cnt_a[:, 1000:1256]
    admissible = ((sig_a[1000:1256] != 0) & (sig_a[1000:1256] != 2147483647)).sum(axis=1)   # rule 5: 0 and M never collide
    assert np.array_equal(blk, blk.T) and np.array_equal(np.diag(blk), admissible)
    assert np.array_equal(cnt_a.max(axis=1), admissible) and admissible.min() >= H - 3

    assert (nv_a == k).all() and (idx_a[:, 0] == np.arange(n)).all() and (dist_a[:, 0] == 0).all()
    assert (np.diff(dist_a, axis=1) >= 0).all()
    for q in rng.choice(n, size=32, replace=False):
        xq = dict(zip(ix[ip[q]:ip[q + 1]].tolist(), dt[ip[q]:ip[q + 1]].astype(np.float64).tolist()))
        nq2 = sum(v * v for v in xq.values())
        for d, got in zip(idx_a[q, 1::9], dist_a[q, 1::9]):
            yi, yv = ix[ip[d]:ip[d + 1]], dt[ip[d]:ip[d + 1]].astype(np.float64)
            dot = sum(xq.get(f, 0.0) * v for f, v in zip(yi.tolist(), yv.tolist()))
            want = np.sqrt(max(0.0, nq2 + float((yv * yv).sum()) - 2.0 * dot))
            assert abs(got - want) <= RTOL * want + 1e-12


def test_config3_full_size_against_the_oracle(cuda, oracle):
    """The headline configuration as a WHOLE against the CPU oracle, nothing sampled: 10 000 cells, H = 800, k = 100,
    exact rerank. Signature matrix byte for byte, a 1024-row block of collision counts against all 10 000 columns, all
    10^6 neighbour indices identical up to distance ties, n_valid identical, distances within 1e-5 relative. The same
    comparison through the reference-facing classes (MinHash.fit -> kneighbors_graph) for the graph structure."""
    cfg, ip, ix, dt, X = cells("s10", 10_000)
    n, H, k = 10_000, 800, 100
    kw = dict(n_neighbors=k, fast=False, **REF_KWARGS)
    ho = oracle.create(number_of_cores=os.cpu_count() or 1, **kw)
    oracle.fit_csr(ho, ip, ix, dt, False)
    o_sig = oracle.signatures(ho, H)
    o_idx, o_dist, o_nv = oracle.kneighbors(ho, k)
    o_cnt = oracle.collision_counts(ho, 4096, 5120)
    o_gip, o_gix, o_gdt = oracle.kneighbors_graph(ho, k)
    oracle.destroy(ho)

    h = cuda.create(**kw)
    cuda.set_feature_range(h, cfg.chroms.nbins ** 2)
    cuda.fit_csr(h, ip, ix, dt, False)
    g_sig = cuda.signatures(h, H)
    assert g_sig.tobytes() == o_sig.tobytes()                      # memcmp of the 32 MB signature matrix
    assert np.array_equal(cuda.collision_counts(h, 4096, 5120), o_cnt)
    g_idx, g_dist, g_nv = cuda.kneighbors(h, k)
    cuda.destroy(h)
    assert np.array_equal(g_nv, o_nv)
    assert_neighbours_equal_up_to_ties(g_idx, g_dist, g_nv, o_idx, o_dist, o_nv, RTOL)
    # integer counts: both sides compute sqrt of the exact integer squared distance -> identical fp32 bits
    assert np.array_equal(g_dist, o_dist)

    from schicexplorer_b200 import MinHash
    mh = MinHash(device=0, **kw)
    g = mh.fit(X).kneighbors_graph(mode="distance")
    mh.close()
    assert g.shape == (n, n) and np.array_equal(g.indptr, o_gip)
    same_cols = np.array_equal(g.indices, o_gix)
    if not same_cols:            # ties at the k-th distance may swap equidistant columns
        assert (g.indices != o_gix).mean() < 1e-4
    np.testing.assert_allclose(np.sort(g.data), np.sort(o_gdt.astype(np.float64)), rtol=RTOL)
