"""GPU suite (-m gpu): the SURVEY 8(f) "next" rows built on the hot path's kernels.

Row 3 - exact kNN (scHicCluster -drm knn, scHicCluster.py:179-180): K4 run as a brute-force sparse euclidean kNN,
checked against the oracle AND against vectors produced by the reference's own scikit-learn call
(tests/golden/knn_golden.npz) - this row is pinned.
Row 4 - hyperopt fast path (scHicClusterMinHashHyperopt.py:119-148): one resident fit, the number of hash functions
switched per trial; every switch must equal a fresh fit with that hash count."""
import os

import numpy as np
import pytest
from scipy.sparse import csr_matrix, vstack

from conftest import GOLDEN, REF_KWARGS, Fitted
from schicexplorer_b200 import MinHash, NearestNeighbors, synth
from test_exact_knn import assert_graph_equal_up_to_ties, exact_graph, synthetic150
from test_gpu_parity import RTOL, assert_neighbours_equal_up_to_ties, synth_csr

pytestmark = pytest.mark.gpu


# ---- row 3: exact kNN ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,k", [("f20", 5), ("f20", 19), ("s150", 10)])
def test_exact_knn_matches_reference_golden_and_oracle(cuda, oracle, fixture20, tag, k):
    g = np.load(os.path.join(GOLDEN, "knn_golden.npz"))
    X = fixture20[0] if tag == "f20" else synthetic150()
    got = exact_graph(cuda, X, k)
    assert_graph_equal_up_to_ties(got, (g[f"{tag}_k{k}_indptr"], g[f"{tag}_k{k}_indices"], g[f"{tag}_k{k}_data"]))
    assert_graph_equal_up_to_ties(got, exact_graph(oracle, X, k))


@pytest.mark.parametrize("mode", ["window", "hash"])
def test_exact_knn_lists_vs_oracle(cuda, oracle, monkeypatch, mode):
    """Neighbour lists (not only the graph), with and without the query row, on clustered cells plus empty, tiny and
    duplicated rows; k from 1 to n - 1; both K4 kernels."""
    monkeypatch.setenv("SCHIC_K4_MODE", mode)
    Xs = synth_csr(180, seed=23)
    ncols = Xs.shape[1]
    one = csr_matrix(([2.0], ([0], [17])), shape=(1, ncols))
    X = vstack([csr_matrix((1, ncols)), Xs[:90], one, Xs[40:43], Xs[90:], csr_matrix((1, ncols))]).tocsr()
    n = X.shape[0]
    for include_self in (False, True):
        for k in (1, 25, n - 1):
            hg = cuda.create(n_neighbors=k, number_of_hash_functions=1, fast=False)
            ho = oracle.create(n_neighbors=k, number_of_hash_functions=1, fast=False)
            for api, h in ((cuda, hg), (oracle, ho)):
                api.fit_csr(h, X.indptr, X.indices, X.data.astype(np.float32), False)
            a = cuda.kneighbors_exact(hg, k, include_self)
            b = oracle.kneighbors_exact(ho, k, include_self)
            assert_neighbours_equal_up_to_ties(*a, *b, RTOL)
            assert (a[2] == k).all()
            if not include_self:
                assert (a[0] != np.arange(n)[:, None]).all()
            cuda.destroy(hg); oracle.destroy(ho)


def test_exact_knn_many_chunks(cuda, oracle):
    """2 500 rows: three K4 chunks per query and a merge over 3 * min(k + 1, 1024) partial results."""
    rng = np.random.default_rng(3)
    ncols = 1 << 20
    rows, cols = [], []
    pools = [rng.choice(ncols, size=1500, replace=False) for _ in range(4)]
    for r in range(2500):
        c = np.unique(rng.choice(pools[r % 4], size=60, replace=False))
        rows.append(np.full(len(c), r)); cols.append(c)
    cc = np.concatenate(cols)
    X = csr_matrix((rng.integers(1, 20, size=len(cc)).astype(np.float64), (np.concatenate(rows), cc)), shape=(2500, ncols))
    for k in (30, 1100):
        assert_graph_equal_up_to_ties(exact_graph(cuda, X, k), exact_graph(oracle, X, k))


def test_exact_knn_against_an_external_database(cuda, oracle):
    """The sharded form on one GPU: two handles hold half of the rows each, both are pointed at the whole CSR
    (schic_mh_set_database_device) and answer for their own rows - the query row is excluded by its GLOBAL index."""
    import torch
    X = synth_csr(170, seed=37)
    dev = torch.device("cuda:0")
    k = 9
    ip = torch.from_numpy(X.indptr.astype(np.int64)).to(dev)
    ix = torch.from_numpy(X.indices.astype(np.int32)).to(dev)
    dt = torch.from_numpy(X.data.astype(np.float32)).to(dev)
    ho = oracle.create(n_neighbors=k, number_of_hash_functions=8, fast=False)
    oracle.fit_csr(ho, X.indptr, X.indices, X.data.astype(np.float32), False)
    ref = {inc: oracle.kneighbors_exact(ho, k, inc) for inc in (False, True)}
    oracle.destroy(ho)
    sig_all = torch.zeros((170, 8), dtype=torch.int32, device=dev)       # not used by the exact search
    for a, b in ((0, 60), (60, 170)):
        h = cuda.create(n_neighbors=k, number_of_hash_functions=8, fast=False, device=0)
        S = X[a:b]
        cuda.fit_csr(h, S.indptr, S.indices, S.data.astype(np.float32), False)
        cuda.set_database_device(h, 170, a, sig_all.data_ptr(), ip.data_ptr(), ix.data_ptr(), dt.data_ptr())
        out = (torch.empty((b - a, k), dtype=torch.int32, device=dev), torch.empty((b - a, k), dtype=torch.float32, device=dev),
               torch.empty(b - a, dtype=torch.int32, device=dev))
        for inc in (False, True):
            cuda.kneighbors_exact_device(h, k, inc, *(t.data_ptr() for t in out))
            cuda.synchronize(h)
            got = [t.cpu().numpy() for t in out]
            assert_neighbours_equal_up_to_ties(*got, ref[inc][0][a:b], ref[inc][1][a:b], ref[inc][2][a:b], RTOL)
            if not inc:
                assert (got[0] != np.arange(a, b)[:, None]).all()
        cuda.destroy(h)


def test_nearest_neighbors_class_on_the_gpu(cuda):
    from sklearn.neighbors import NearestNeighbors as SkNN
    X = synthetic150()
    nn = NearestNeighbors(n_neighbors=12, algorithm="ball_tree").fit(X)
    assert nn._api.backend == "cuda"
    dist, idx = nn.kneighbors()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sd, si = SkNN(n_neighbors=12, algorithm="ball_tree").fit(X).kneighbors()
    np.testing.assert_allclose(dist, sd, rtol=RTOL)
    same = idx == si
    assert same.mean() > 0.99 and np.allclose(dist[~same], sd[~same], rtol=RTOL)     # only ties may swap
    g = nn.kneighbors_graph(mode="distance")
    assert g.shape == (150, 150) and (np.diff(g.indptr) == 12).all()
    nn.close()


def test_exact_knn_s10_shape_sampled(cuda):
    """1 500 full-length S10-shaped cells (14 k non-zeros each): 2.2e6 pairs through the windowed kernel; sampled lists
    recomputed directly from the sparse rows in float64."""
    cfg = synth.SynthConfig(**{**synth.CONFIGS["s10"].__dict__, "n_cells": 1500})
    ip, ix, dt, _ = synth.generate_csr(cfg, 0, 1500, workers=min(16, os.cpu_count() or 1))
    X = synth.to_scipy(cfg, ip, ix, dt)
    h = cuda.create(n_neighbors=20, number_of_hash_functions=1, fast=False)
    cuda.set_feature_range(h, X.shape[1])
    cuda.fit_csr(h, ip, ix, dt, False)
    idx, dist, nv = cuda.kneighbors_exact(h, 20, include_self=False)
    cuda.destroy(h)
    assert (nv == 20).all() and (np.diff(dist, axis=1) >= 0).all()
    rng = np.random.default_rng(1)
    Xd = X.astype(np.float64)
    sq = np.asarray(Xd.multiply(Xd).sum(axis=1)).ravel()
    for q in rng.choice(1500, size=12, replace=False):
        d2 = sq[q] + sq - 2.0 * np.asarray((Xd @ Xd[q].T).todense()).ravel()
        d = np.sqrt(np.maximum(d2, 0.0)); d[q] = np.inf
        order = np.lexsort((np.arange(1500), d))[:20]
        np.testing.assert_allclose(dist[q], d[order], rtol=RTOL)
        assert set(idx[q].tolist()) == set(order.tolist())


# ---- row 4: hyperopt fast path ------------------------------------------------------------------------------------
def test_active_hash_prefix_equals_fresh_fit(cuda, oracle):
    """One resident fit at the sweep's largest hash count; every switched-to hash count gives the signatures, counts
    and neighbour lists of a fresh fit with that count (the oracle's)."""
    X = synth_csr(260, seed=29)
    kw = dict(REF_KWARGS, number_of_hash_functions=1200)
    hg = cuda.create(n_neighbors=30, fast=False, **kw)
    cuda.fit_csr(hg, X.indptr, X.indices, X.data.astype(np.float32), False)
    for H in (400, 1200, 37, 800, 800, 1):
        cuda.set_active_hashes(hg, H)
        okw = dict(REF_KWARGS, number_of_hash_functions=H, minimal_blocks_in_common=min(100, H))
        with Fitted(oracle, X, fast=False, n_neighbors=30, **okw) as o:
            assert np.array_equal(cuda.signatures(hg, H), o.sig())
            if H >= 100:     # handle keeps minimal_blocks_in_common = 100 (as the reference's CLI does across -nh)
                assert np.array_equal(cuda.collision_counts(hg), oracle.collision_counts(o.h))
                assert_neighbours_equal_up_to_ties(*cuda.kneighbors(hg, 30), *oracle.kneighbors(o.h, 30), RTOL)
                a, b = cuda.kneighbors(hg, 30, fast=1), oracle.kneighbors(o.h, 30, fast=1)
                assert all(np.array_equal(x, y) for x, y in zip(a, b))
    with pytest.raises(Exception):
        cuda.set_active_hashes(hg, 1201)
    # a partial_fit switches back to the created hash count and appends to the full signatures
    cuda.set_active_hashes(hg, 300)
    Y = synth_csr(40, seed=30)
    cuda.fit_csr(hg, Y.indptr, Y.indices, Y.data.astype(np.float32), True)
    XY = vstack([X, Y]).tocsr()
    with Fitted(oracle, XY, fast=False, n_neighbors=30, **kw) as o:
        assert np.array_equal(cuda.signatures(hg, 1200), o.sig())
        cuda.set_active_hashes(hg, 600)
    with Fitted(oracle, XY, fast=True, n_neighbors=30, **dict(kw, number_of_hash_functions=600)) as o:
        a, b = cuda.kneighbors(hg, 30, fast=1), oracle.kneighbors(o.h, 30, fast=1)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    cuda.destroy(hg)


def test_minhash_class_hash_sweep(cuda, oracle):
    """MinHash.set_number_of_hash_functions: the sweep of scHicClusterMinHashHyperopt without re-fitting."""
    X = synth_csr(150, seed=33)
    mh = MinHash(n_neighbors=150, fast=True, **dict(REF_KWARGS, number_of_hash_functions=1000))
    mh.fit(X)
    for H in (200, 600, 1000):
        g = mh.set_number_of_hash_functions(H).kneighbors_graph(mode="distance")
        with Fitted(oracle, X, fast=True, n_neighbors=150, **dict(REF_KWARGS, number_of_hash_functions=H)) as o:
            ip, ix, dt = oracle.kneighbors_graph(o.h, 150)
        assert np.array_equal(g.indptr, ip) and np.array_equal(g.indices, ix)
        np.testing.assert_array_equal(g.data.astype(np.float32), dt)
    with pytest.raises(ValueError):
        mh.set_number_of_hash_functions(1001)
    mh.fit(X)
    assert mh.number_of_hash_functions == 1000 and mh.signatures().shape == (150, 1000)
    mh.close()


def test_hyperopt_tool_on_gpu(cuda, oracle, fixture20, tmp_path):
    """scHicClusterMinHashHyperopt on the CUDA back end: each CSR hashed once, every trial's graph == a fresh fit's."""
    from schicexplorer_b200 import loader, scHicClusterMinHashHyperopt as hyper
    matrix = os.path.join(GOLDEN, "test_matrix.scool")
    res = hyper.ResidentCells(matrix, 1000, 1000)
    for intra, H in [(False, 400), (True, 1000), (False, 1000), (True, 250), (False, 400)]:
        g = res.graph(intra, H)
        X, _ = loader.cells_to_csr(res.cells, res.chroms, intra_only=intra)
        with Fitted(oracle, X, fast=True, n_neighbors=20, **dict(REF_KWARGS, number_of_hash_functions=H)) as fresh:
            ip, ix, dt = oracle.kneighbors_graph(fresh.h, 20)
        assert np.array_equal(g.indptr, ip) and np.array_equal(g.indices, ix)
        np.testing.assert_array_equal(g.data.astype(np.float32), dt)
    assert res.fits == 2 and res.minhash(False)._api.backend == "cuda"
    res.close()
    colors = tmp_path / "cell_types.txt"
    colors.write_text("".join(f"{n}\t{'ABC'[i % 3]}\n" for i, n in enumerate(res.names)))
    out = tmp_path / "best.txt"
    best = hyper.main(f"-m {matrix} -c {colors} -o {out} --runs 4 -noh 200 900 200 -noc 2 5 -nop 5 16 5 -me random".split())
    assert out.read_text().startswith("# Created by scHiCExplorer") and best["numberOfHashFunctions"] in (200, 400, 600, 800)


# ---- row 2: dense-graph PCA -------------------------------------------------------------------------------------------
def _pca_reference(X, dim):
    from sklearn.decomposition import PCA as SkPCA
    full = SkPCA(n_components=min(X.shape) - 1).fit(X)
    return full.transform(X)[:, :dim], full.singular_values_[:dim], full.components_[:dim]


def _close(got, want, rtol=1e-9):
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= rtol * np.abs(want).max()


def test_pca_fixture_graph_vs_sklearn_and_oracle(cuda, oracle, fixture20):
    """The graph scHicClusterMinHash builds on the reference's 20 cells -> PCA(n-1 components)[:, :dim]: device ==
    scikit-learn (the reference's call) == numpy oracle; CSR input (densified on the device) == dense input."""
    from oracle import pca_np
    with Fitted(cuda, fixture20[0], fast=True, n_neighbors=20) as f:
        ip, ix, dt = cuda.kneighbors_graph(f.h, 20)
    G = csr_matrix((dt, ix, ip), shape=(20, 20))
    X = np.asarray(G.todense(), dtype=np.float64)
    for dim in (2, 19):
        emb, sv, comp = _pca_reference(X, dim)
        got, gsv, gcomp = cuda.pca_fit_transform(X, dim, want_components=True)
        _close(got, emb); _close(gcomp, comp)
        np.testing.assert_allclose(gsv, sv, rtol=1e-10, atol=1e-12 * sv[0])
        got2, gsv2, _ = cuda.pca_fit_transform(G, dim)
        _close(got2, emb)
        o_emb, o_sv, _ = pca_np.pca_fit_transform(X, dim)
        _close(got, o_emb)
    g = np.load(os.path.join(GOLDEN, "pca_golden.npz"))          # committed sklearn output (make_pca_golden.py)
    got, gsv, gcomp = cuda.pca_fit_transform(G, 19, want_components=True)
    _close(got, g["embedding"]); _close(gcomp, g["components"])
    np.testing.assert_allclose(cuda.last_pca_mean, g["mean"], rtol=1e-12, atol=1e-15)
    st = cuda.pca_stage_ms()
    assert st["total"] > 0 and abs(sum(v for k, v in st.items() if k != "total") - st["total"]) < 1e-3 * st["total"] + 1e-3


@pytest.mark.parametrize("n,dim,seed", [(1, 1, 0), (2, 1, 0), (3, 2, 0), (65, 64, 0), (150, 100, 1), (700, 100, 2), (1500, 60, 3)])
def test_pca_graph_like_matrices(cuda, n, dim, seed):
    from test_pca import graph_like
    if n < 4:
        X = np.random.default_rng(seed).standard_normal((n + 3, n))      # tiny feature counts
        want = None
    else:
        X = graph_like(n, seed)
    got, gsv, comp = cuda.pca_fit_transform(X, dim, want_components=True)
    if n < 4:
        from sklearn.decomposition import PCA as SkPCA
        full = SkPCA(n_components=dim).fit(X)
        _close(got, full.transform(X)); _close(comp, full.components_)
        return
    emb, sv, rcomp = _pca_reference(X, dim)
    _close(got, emb); _close(comp, rcomp)
    np.testing.assert_allclose(gsv, sv, rtol=1e-10, atol=1e-12 * sv[0])


def test_pca_full_square_tridiagonal_pass_agrees(cuda, monkeypatch):
    """SCHIC_PCA_TRIDIAG=full selects the first (full-square) trailing pass; the default walks the lower triangle only."""
    from test_pca import graph_like
    X = graph_like(333, 11)
    a = cuda.pca_fit_transform(X, 80, want_components=True)
    monkeypatch.setenv("SCHIC_PCA_TRIDIAG", "full")
    b = cuda.pca_fit_transform(X, 80, want_components=True)
    _close(a[0], b[0]); _close(a[2], b[2])
    np.testing.assert_allclose(a[1], b[1], rtol=1e-12)
    emb, _, _ = _pca_reference(X, 80)
    _close(b[0], emb)


def test_pca_rectangular_scrub_and_errors(cuda):
    from test_pca import graph_like
    rng = np.random.default_rng(3)
    X = rng.standard_normal((120, 70))
    emb, _, _ = _pca_reference(X, 30)
    _close(cuda.pca_fit_transform(X, 30)[0], emb)
    X = rng.standard_normal((70, 120))
    emb, _, _ = _pca_reference(X, 30)
    _close(cuda.pca_fit_transform(X, 30)[0], emb)
    Y = graph_like(90, 4)
    Y[[5, 17, 60]] = 0.0; Y[:, [5, 17, 60]] = 0.0
    Z = Y.copy(); Z[3, 8] = np.inf; Z[4, 9] = np.nan
    Y[3, 8] = 0.0; Y[4, 9] = 0.0
    emb, _, _ = _pca_reference(Y, 25)
    _close(cuda.pca_fit_transform(Z, 25)[0], emb)
    Zs = csr_matrix(Z.astype(np.float32))
    emb32, _, _ = _pca_reference(np.nan_to_num(np.asarray(Zs.todense(), dtype=np.float64), nan=0.0, posinf=0.0, neginf=0.0), 25)
    _close(cuda.pca_fit_transform(Zs, 25)[0], emb32)
    with pytest.raises(Exception):
        cuda.pca_fit_transform(Y, 91)
    with pytest.raises(Exception):
        cuda.pca_fit_transform(Y, 0)


def test_pca_class_and_clustering_post_step(cuda, fixture20):
    """schicexplorer_b200.PCA mirrors the sklearn surface the tools use; MinHashClustering's post-step runs on it."""
    from sklearn.decomposition import PCA as SkPCA
    from schicexplorer_b200 import MinHashClustering
    from schicexplorer_b200.pca import PCA
    from test_pca import graph_like
    X = graph_like(200, 7)
    ref = SkPCA(n_components=199).fit(X)
    p = PCA(n_components=199, keep=40)
    emb = p.fit_transform(X)
    _close(emb, ref.transform(X)[:, :40])
    _close(p.components_, ref.components_[:40])
    np.testing.assert_allclose(p.singular_values_, ref.singular_values_[:40], rtol=1e-10)
    np.testing.assert_allclose(p.explained_variance_, ref.explained_variance_[:40], rtol=1e-9)
    np.testing.assert_allclose(p.mean_, ref.mean_, rtol=1e-12)
    G = csr_matrix(X.astype(np.float32))         # graphs cross the C ABI as float32 (as kneighbors_graph returns them)
    ref32 = SkPCA(n_components=199).fit_transform(np.asarray(G.todense(), dtype=np.float64))
    out = MinHashClustering._post_process(G, True, 30, False, None)
    _close(out, ref32[:, :30])


def test_pca_degenerate_inputs_stay_finite(cuda):
    """All-zero graph (no cell has a neighbour) and identical rows: an all-zero, finite embedding, as scikit-learn gives."""
    for X in (np.zeros((12, 12)), np.tile(np.arange(9.0), (9, 1))):
        got, sv, comp = cuda.pca_fit_transform(X, 4, want_components=True)
        assert np.isfinite(got).all() and np.isfinite(comp).all() and np.isfinite(sv).all()
        assert np.abs(got).max() < 1e-9 and sv.max() < 1e-9


def test_exact_knn_tiny_inputs(cuda, oracle):
    """One, two and three rows (a single row has no neighbour once it is excluded itself), an empty row among them."""
    rows = [csr_matrix(([1.0, 2.0, 5.0], ([0, 0, 0], [3, 10, 77])), shape=(1, 100)),
            csr_matrix((1, 100)),
            csr_matrix(([4.0, 2.0], ([0, 0], [10, 50])), shape=(1, 100))]
    for n in (1, 2, 3):
        X = vstack(rows[:n]).tocsr()
        for include_self in (False, True):
            for k in (1, 3):
                res = []
                for api in (cuda, oracle):
                    h = api.create(n_neighbors=k, number_of_hash_functions=4, fast=False)
                    api.fit_csr(h, X.indptr, X.indices, X.data.astype(np.float32), False)
                    res.append(api.kneighbors_exact(h, k, include_self))
                    api.destroy(h)
                assert_neighbours_equal_up_to_ties(*res[0], *res[1], RTOL)
