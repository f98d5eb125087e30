"""CPU suite: the dense-graph PCA post-step -- the numpy oracle (oracle/pca_np.py, the same numerical
steps as csrc/pca_dense.cu) held to scikit-learn's PCA, the reference's own call
(scHicClusterMinHash.py:358-366: PCA(n_components=min(shape) - 1).fit_transform(dense)[:, :dim]). This pins the row."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix
from sklearn.decomposition import PCA as SkPCA

from conftest import Fitted
from oracle import pca_np

RTOL = 1e-9


def reference_embedding(X, dim):
    full = SkPCA(n_components=min(X.shape) - 1).fit(X)
    return full.transform(X)[:, :dim], full.singular_values_[:dim], full.components_[:dim]


def assert_embedding_close(got, want, rtol=RTOL):
    scale = np.abs(want).max()
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= rtol * scale


def graph_like(n, seed, density=0.6):
    """A dense matrix with the look of a kNN distance graph: block structure + noise, zeros for missing edges."""
    rng = np.random.default_rng(seed)
    z = rng.integers(0, 4, size=n)
    base = 0.55 + 0.3 * (z[:, None] != z[None, :]) + 0.05 * rng.standard_normal((n, n))
    base[rng.random((n, n)) > density] = 0.0
    np.fill_diagonal(base, 0.0)
    return np.abs(base)


@pytest.fixture(scope="module")
def fixture_graph(oracle, fixture20):
    with Fitted(oracle, fixture20[0], fast=True, n_neighbors=20) as f:
        ip, ix, dt = oracle.kneighbors_graph(f.h, 20)
    return csr_matrix((dt, ix, ip), shape=(20, 20))


def test_oracle_on_the_reference_fixture_graph(fixture_graph):
    X = np.asarray(fixture_graph.todense(), dtype=np.float64)
    for dim in (2, 10, 19):
        emb, sv, comp = reference_embedding(X, dim)
        got, gsv, gcomp = pca_np.pca_fit_transform(X, dim)
        assert_embedding_close(got, emb)
        np.testing.assert_allclose(gsv, sv, rtol=1e-10, atol=1e-12 * sv[0])
        assert_embedding_close(gcomp, comp)


def test_oracle_against_committed_golden_vectors(fixture_graph):
    """tests/golden/pca_golden.npz (make_pca_golden.py): sklearn's output on the fixture graph, committed."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "pca_golden.npz"))
    assert np.array_equal(fixture_graph.indptr, g["graph_indptr"]) and np.array_equal(fixture_graph.indices, g["graph_indices"])
    assert np.array_equal(fixture_graph.data, g["graph_data"])
    X = np.asarray(fixture_graph.todense(), dtype=np.float64)
    got, sv, comp = pca_np.pca_fit_transform(X, 19)
    assert_embedding_close(got, g["embedding"])
    assert_embedding_close(comp, g["components"])
    np.testing.assert_allclose(sv, g["singular_values"], rtol=1e-10, atol=1e-12 * sv[0])


@pytest.mark.parametrize("n,dim,seed", [(64, 63, 0), (150, 100, 1), (300, 40, 2)])
def test_oracle_on_graph_like_matrices(n, dim, seed):
    X = graph_like(n, seed)
    emb, sv, _ = reference_embedding(X, dim)
    got, gsv, _ = pca_np.pca_fit_transform(X, dim)
    assert_embedding_close(got, emb)
    np.testing.assert_allclose(gsv, sv, rtol=1e-10, atol=1e-12 * sv[0])


def test_oracle_rectangular_and_rank_deficient():
    rng = np.random.default_rng(3)
    X = rng.standard_normal((120, 70))
    emb, sv, _ = reference_embedding(X, 30)
    got, gsv, _ = pca_np.pca_fit_transform(X, 30)
    assert_embedding_close(got, emb)
    # rows without any neighbour (all-zero rows and columns) and a non-finite entry (scrubbed to 0 as the reference does)
    Y = graph_like(90, 4)
    Y[[5, 17, 60]] = 0.0
    Y[:, [5, 17, 60]] = 0.0
    Z = Y.copy()
    Z[3, 8] = np.inf
    Y[3, 8] = 0.0
    emb, sv, _ = reference_embedding(Y, 25)
    got, gsv, _ = pca_np.pca_fit_transform(Z, 25)
    assert_embedding_close(got, emb)


def test_oracle_stages():
    """The building blocks against numpy.linalg: tridiagonal form has the same spectrum, bisection finds it."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((80, 80))
    C = A.T @ A
    d, e, V, tau = pca_np.tridiagonalize(C)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    ev = np.linalg.eigvalsh(C)[::-1]
    np.testing.assert_allclose(np.linalg.eigvalsh(T)[::-1], ev, rtol=1e-10, atol=1e-10 * ev[0])
    np.testing.assert_allclose(pca_np.top_eigenvalues(d, e, 30), ev[:30], rtol=1e-10, atol=1e-12 * ev[0])
    X = pca_np.tridiag_vectors(d, e, ev[:30])
    np.testing.assert_allclose(X.T @ X, np.eye(30), atol=1e-10)
    np.testing.assert_allclose(T @ X, X * ev[:30], atol=1e-8 * ev[0])
    Z = pca_np.back_transform(V, tau, X)
    np.testing.assert_allclose(C @ Z, Z * ev[:30], atol=1e-8 * ev[0])


def test_oracle_degenerate_inputs_stay_finite():
    """No neighbours anywhere (all-zero graph) or identical rows: sklearn returns an all-zero embedding; so must we."""
    for X in (np.zeros((12, 12)), np.tile(np.arange(9.0), (9, 1))):
        want = SkPCA(n_components=min(X.shape) - 1).fit_transform(X)[:, :4]
        got, sv, comp = pca_np.pca_fit_transform(X, 4)
        assert np.isfinite(got).all() and np.isfinite(comp).all() and np.isfinite(sv).all()
        assert np.abs(got).max() < 1e-9 and np.abs(want).max() < 1e-9 and sv.max() < 1e-9


@pytest.mark.parametrize("n,nb", [(1, 4), (2, 4), (7, 3), (40, 8), (97, 32), (130, 64)])
def test_blocked_tridiagonalisation_equals_unblocked(n, nb):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n + 5, n))
    C = A.T @ A
    d0, e0, V0, t0 = pca_np.tridiagonalize(C)
    d1, e1, V1, t1 = pca_np.tridiagonalize_blocked(C, nb)
    scale = np.abs(d0).max()
    np.testing.assert_allclose(d1, d0, atol=1e-11 * scale)
    np.testing.assert_allclose(e1, e0, atol=1e-11 * scale)
    np.testing.assert_allclose(t1, t0, atol=1e-10)
    np.testing.assert_allclose(V1, V0, atol=1e-9)
