"""CPU suite: the exact-kNN sibling (scHicCluster.py:179-180).

This row IS pinned to the reference's arithmetic: the golden vectors in tests/golden/knn_golden.npz were produced by the
very scikit-learn call the reference makes (tests/golden/make_knn_golden.py), and scikit-learn is importable in the
test image, so the oracle is also checked against a live sklearn run on seeded inputs. The GPU suite then holds the
CUDA path to the oracle and to the same golden vectors (tests/test_gpu_next.py)."""
import os
import warnings

import numpy as np
import pytest
from scipy.sparse import random as sprandom

from conftest import GOLDEN
from schicexplorer_b200 import synth
from schicexplorer_b200.knn import NearestNeighbors

RTOL = 1e-5


def graph_rows(indptr, indices, data):
    return [(indices[indptr[r]:indptr[r + 1]], data[indptr[r]:indptr[r + 1]]) for r in range(len(indptr) - 1)]


def assert_graph_equal_up_to_ties(got, want, rtol=RTOL):
    """Both graphs as (indptr, indices, data) with rows sorted by column. Rows must hold the same number of
    neighbours; where the neighbour sets differ the distances involved must tie with the row's k-th distance."""
    assert np.array_equal(got[0], want[0])
    for (gi, gd), (wi, wd) in zip(graph_rows(*got), graph_rows(*want)):
        if np.array_equal(gi, wi):
            np.testing.assert_allclose(gd, wd, rtol=rtol, atol=1e-9)
            continue
        kth = wd.max()
        only_g, only_w = ~np.isin(gi, wi), ~np.isin(wi, gi)
        np.testing.assert_allclose(gd[only_g], kth, rtol=rtol)
        np.testing.assert_allclose(wd[only_w], kth, rtol=rtol)
        np.testing.assert_allclose(gd[~only_g], wd[~only_w], rtol=rtol, atol=1e-9)


def synthetic150():
    cfg = synth.small_config(150, config_id=41)
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=1)
    return synth.to_scipy(cfg, ip, ix, dt)


def sklearn_graph(X, k):
    from sklearn.neighbors import NearestNeighbors as SkNN
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        g = SkNN(n_neighbors=k, algorithm="ball_tree").fit(X).kneighbors_graph(mode="distance").tocsr()
    g.sort_indices()
    return g.indptr.astype(np.int64), g.indices.astype(np.int32), g.data


def exact_graph(api, X, k, include_self=False):
    h = api.create(n_neighbors=k, number_of_hash_functions=4, fast=False)
    try:
        api.fit_csr(h, X.indptr, X.indices, X.data.astype(np.float32), False)
        return api.kneighbors_exact_graph(h, k, include_self)
    finally:
        api.destroy(h)


@pytest.mark.parametrize("tag,k", [("f20", 5), ("f20", 19), ("s150", 10)])
def test_oracle_matches_the_reference_call_golden(oracle, fixture20, tag, k):
    g = np.load(os.path.join(GOLDEN, "knn_golden.npz"))
    X = fixture20[0] if tag == "f20" else synthetic150()
    want = (g[f"{tag}_k{k}_indptr"], g[f"{tag}_k{k}_indices"], g[f"{tag}_k{k}_data"])
    assert_graph_equal_up_to_ties(exact_graph(oracle, X, k), want)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_live_sklearn(oracle, seed):
    rng = np.random.default_rng(seed)
    X = sprandom(60, 5000, density=0.02, random_state=seed, format="csr", dtype=np.float64)
    X.data = rng.integers(1, 40, size=X.nnz).astype(np.float64)
    for k in (1, 7, 59):
        assert_graph_equal_up_to_ties(exact_graph(oracle, X, k), sklearn_graph(X, k))


def test_include_self_and_padding(oracle):
    X = synthetic150()[:30]
    h = oracle.create(n_neighbors=5, number_of_hash_functions=4, fast=False)
    oracle.fit_csr(h, X.indptr, X.indices, X.data.astype(np.float32), False)
    idx, dist, nv = oracle.kneighbors_exact(h, 40, include_self=True)     # k > n: padded
    assert (nv == 30).all() and (idx[:, 0] == np.arange(30)).all() and (dist[:, 0] == 0).all()
    assert (idx[:, 30:] == -1).all() and np.isinf(dist[:, 30:]).all()
    idx2, dist2, nv2 = oracle.kneighbors_exact(h, 29, include_self=False)
    assert (nv2 == 29).all() and np.array_equal(idx2, idx[:, 1:30]) and np.array_equal(dist2, dist[:, 1:30])
    oracle.destroy(h)


def test_class_mirrors_sklearn_surface(oracle, monkeypatch):
    """The host-side mirror on the oracle back end (host logic only; the GPU suite runs it on the CUDA library)."""
    monkeypatch.setattr(NearestNeighbors, "_api_factory", staticmethod(lambda: oracle))
    X = synthetic150()
    nn = NearestNeighbors(n_neighbors=10, algorithm="ball_tree", n_jobs=4).fit(X)
    g = nn.kneighbors_graph(mode="distance")
    assert g.shape == (150, 150) and g.dtype == np.float64 and (np.diff(g.indptr) == 10).all()
    gs = g.copy(); gs.sort_indices()
    assert_graph_equal_up_to_ties((gs.indptr.astype(np.int64), gs.indices, gs.data), sklearn_graph(X, 10))
    c = nn.kneighbors_graph(n_neighbors=3)
    assert c.nnz == 450 and (c.data == 1).all()
    dist, idx = nn.kneighbors()
    assert dist.shape == idx.shape == (150, 10) and (np.diff(dist, axis=1) >= 0).all()
    assert (idx != np.arange(150)[:, None]).all()
    with pytest.raises(ValueError):
        nn.kneighbors_graph(n_neighbors=150)          # sklearn: n_neighbors <= n_samples_fit - 1 for X=None
    with pytest.raises(ValueError):
        nn.kneighbors_graph(mode="similarity")
    with pytest.raises(NotImplementedError):
        nn.kneighbors(X)
    with pytest.raises(ValueError):
        NearestNeighbors(metric="cosine")
    nn.close()
