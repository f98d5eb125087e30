"""CPU suite: the oracle (oracle/oracle.cpp) against the committed golden vectors, against the
independent numpy restatement, and the size-independent properties of the path."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix, vstack

from conftest import REF_KWARGS, Fitted
from oracle import oracle_np
from schicexplorer_b200 import synth


@pytest.mark.parametrize("width", [32, 64])
def test_signatures_and_counts_match_golden(oracle, fixture20, golden20, width):
    X = fixture20[0]
    with Fitted(oracle, X, hash_width=width) as f:
        assert np.array_equal(f.sig(), golden20[f"sig{width}"])
        assert np.array_equal(oracle.collision_counts(f.h), golden20[f"cnt{width}"])


def test_fixture_statistics_match_survey(golden20):
    # SURVEY.md 8(c) sanity probe: counts min 21 / median 122.5 / max 251; 142 of 190 pairs >= 100
    c = golden20["cnt32"][np.triu_indices(20, 1)]
    assert (c.min(), float(np.median(c)), c.max(), int((c >= 100).sum())) == (21, 122.5, 251, 142)
    assert not (golden20["sig32"] == 0).any() and not (golden20["sig32"] == 2147483647).any()


@pytest.mark.parametrize("fast", [True, False])
def test_kneighbors_match_golden(oracle, fixture20, golden20, fast):
    X = fixture20[0]
    tag = "fast" if fast else "exact"
    with Fitted(oracle, X, fast=fast) as f:
        idx, dist, nv = oracle.kneighbors(f.h, 20)
    assert np.array_equal(nv, golden20[f"nv_{tag}"])
    assert np.array_equal(idx, golden20[f"idx_{tag}"])
    fin = np.isfinite(dist)
    assert np.array_equal(fin, np.isfinite(golden20[f"dist_{tag}"]))
    np.testing.assert_allclose(dist[fin], golden20[f"dist_{tag}"][fin], rtol=1e-6)
    assert (idx[:, 0] == np.arange(20)).all() and (dist[:, 0] == 0).all()   # self first, distance 0


def test_wang_known_values():
    # hand-evaluated: h_0(f=0) = wang32(1) mod M ; checked against a scalar python evaluation
    def w(k):
        k &= 0xFFFFFFFF
        k = (~k + (k << 15)) & 0xFFFFFFFF
        k ^= k >> 12
        k = (k + (k << 2)) & 0xFFFFFFFF
        k ^= k >> 4
        k = (k * 2057) & 0xFFFFFFFF
        k ^= k >> 16
        return k
    X = csr_matrix((np.ones(3), ([0, 0, 1], [0, 7, 123456])), shape=(2, 200000))
    sig = oracle_np.signatures(X, 5)
    for j in range(5):
        assert sig[0, j] == min(w(1 * (j + 1)) % 2147483647, w(8 * (j + 1)) % 2147483647)
        assert sig[1, j] == w(123457 * (j + 1)) % 2147483647


def small_csr(n=60, seed=7, **kw):
    cfg = synth.small_config(n, config_id=seed, **kw)
    ip, ix, dt, z = synth.generate_csr(cfg, workers=1)
    return synth.to_scipy(cfg, ip, ix, dt), z


@pytest.mark.parametrize("width", [32, 64])
def test_oracle_vs_numpy_on_synthetic(oracle, width):
    X, _ = small_csr(40, median_nnz=800, min_nnz=50, max_nnz=3000)
    H = 96
    with Fitted(oracle, X, fast=False, number_of_hash_functions=H, minimal_blocks_in_common=5, hash_width=width) as f:
        sig = f.sig()
        assert np.array_equal(sig, oracle_np.signatures(X, H, width=width))
        assert np.array_equal(oracle.collision_counts(f.h).astype(np.int64), oracle_np.collision_counts(sig))
        for fast in (1, 0):
            idx, dist, nv = oracle.kneighbors(f.h, 7, fast)
            i2, d2, n2 = oracle_np.kneighbors(X, sig, 7, fast=bool(fast), min_blocks=5)
            assert np.array_equal(nv, n2) and np.array_equal(idx, i2)
            fin = np.isfinite(d2)
            np.testing.assert_allclose(dist[fin], d2[fin], rtol=1e-6)


def test_empty_rows_and_sentinel(oracle):
    X, _ = small_csr(12, median_nnz=300, min_nnz=20, max_nnz=900)
    empty = csr_matrix((1, X.shape[1]))
    Xe = vstack([X[:5], empty, X[5:], empty]).tocsr()
    with Fitted(oracle, Xe, number_of_hash_functions=64, minimal_blocks_in_common=1) as f:
        sig = f.sig()
        assert (sig[5] == 2147483647).all() and (sig[-1] == 2147483647).all()
        cnt = oracle.collision_counts(f.h)
        assert (cnt[5] == 0).all() and (cnt[:, 5] == 0).all()      # the sentinel never collides
        idx, dist, nv = oracle.kneighbors(f.h, 4)
        assert nv[5] == 0 and (idx[5] == -1).all() and np.isinf(dist[5]).all()


def test_partial_fit_equals_fit(oracle, fixture20):
    X = fixture20[0]
    with Fitted(oracle, X) as a:
        ref_sig, (ri, rd, rn) = a.sig(), oracle.kneighbors(a.h, 20)
    h = oracle.create(n_neighbors=20, fast=True, **REF_KWARGS)
    oracle.fit_csr(h, X[:7].indptr, X[:7].indices, None, False)
    oracle.fit_csr(h, X[7:15].indptr, X[7:15].indices, None, True)
    oracle.fit_csr(h, X[15:].indptr, X[15:].indices, None, True)
    assert oracle.n_fitted(h) == 20
    assert np.array_equal(oracle.signatures(h, 800), ref_sig)
    i, d, n = oracle.kneighbors(h, 20)
    assert np.array_equal(i, ri) and np.array_equal(n, rn) and np.array_equal(d, rd)
    oracle.destroy(h)


def test_nnz_order_invariance(oracle):
    X, _ = small_csr(10, median_nnz=400, min_nnz=50, max_nnz=900)
    rng = np.random.default_rng(0)
    idx = X.indices.copy()
    for r in range(X.shape[0]):
        a, b = X.indptr[r], X.indptr[r + 1]
        idx[a:b] = rng.permutation(idx[a:b])
    h1 = oracle.create(n_neighbors=5, fast=True, number_of_hash_functions=50)
    h2 = oracle.create(n_neighbors=5, fast=True, number_of_hash_functions=50)
    oracle.fit_csr(h1, X.indptr, X.indices, None, False)
    oracle.fit_csr(h2, X.indptr, idx, None, False)
    assert np.array_equal(oracle.signatures(h1, 50), oracle.signatures(h2, 50))
    oracle.destroy(h1)
    oracle.destroy(h2)


def test_graph_rows_sorted_by_column(oracle, fixture20):
    X = fixture20[0]
    with Fitted(oracle, X) as f:
        indptr, indices, data = oracle.kneighbors_graph(f.h, 20)
        idx, dist, nv = oracle.kneighbors(f.h, 20)
    assert np.array_equal(np.diff(indptr), nv)
    for c in range(20):
        cols = indices[indptr[c]:indptr[c + 1]]
        assert (np.diff(cols) > 0).all()
        order = np.argsort(idx[c, :nv[c]])
        assert np.array_equal(cols, idx[c, :nv[c]][order])
        assert np.array_equal(data[indptr[c]:indptr[c + 1]], dist[c, :nv[c]][order])


def test_max_bin_size_prunes_large_buckets(oracle):
    # 6 identical rows -> every bucket has size 6; with max_bin_size=5 nothing collides
    X, _ = small_csr(1, median_nnz=300, min_nnz=100, max_nnz=600)
    Xr = vstack([X] * 6).tocsr()
    with Fitted(oracle, Xr, number_of_hash_functions=32, minimal_blocks_in_common=1, max_bin_size=5) as f:
        assert (oracle.collision_counts(f.h) == 0).all()
    with Fitted(oracle, Xr, number_of_hash_functions=32, minimal_blocks_in_common=1, max_bin_size=7) as f:
        cnt = oracle.collision_counts(f.h)
        assert (cnt == 32).all()
        assert np.array_equal(cnt.astype(np.int64), oracle_np.collision_counts(f.sig(), max_bin_size=7))


def test_error_reporting(oracle):
    from schicexplorer_b200._capi import SchicError
    with pytest.raises(SchicError):
        oracle.create(n_neighbors=5, number_of_hash_functions=0)
    with pytest.raises(SchicError):
        oracle.create(n_neighbors=5, number_of_hash_functions=8, shingle=1, shingle_size=9)   # no table left
    h = oracle.create(n_neighbors=5)
    with pytest.raises(SchicError) as e:
        oracle.kneighbors(h, 5, n_rows=1)
    assert e.value.status == -2   # SCHIC_ERR_NOT_FITTED
    oracle.destroy(h)


@pytest.mark.parametrize("width", [32, 64])
def test_block_combine_shingle(oracle, fixture20, width):
    """SURVEY 8(c) rule 4 (decision point D2, off by default): with shingle=1 the minima of shingle_size consecutive
    hash functions are folded into one key. The C++ oracle against the numpy restatement: keys, counts, neighbours;
    empty rows keep the sentinel; partial_fit == fit; the active-hash prefix is refused."""
    from schicexplorer_b200._capi import SchicError
    X = fixture20[0]
    empty = csr_matrix((1, X.shape[1]))
    Xe = vstack([X[:9], empty, X[9:]]).tocsr()
    H, s = 803, 5                                   # 160 tables, 3 left-over minima unused
    kw = dict(number_of_hash_functions=H, shingle=1, shingle_size=s, minimal_blocks_in_common=3, hash_width=width)
    T = H // s
    raw = oracle_np.signatures(Xe, H, width=width)
    keys = oracle_np.combine(raw, s, width=width, empty_rows=np.diff(Xe.indptr) == 0)
    assert keys.shape == (21, T) and (keys[9] == 2147483647).all()
    h = oracle.create(n_neighbors=20, fast=False, **kw)
    oracle.fit_csr(h, Xe.indptr, Xe.indices, Xe.data.astype(np.float32), False)
    assert np.array_equal(oracle.signatures(h, T), keys)
    assert np.array_equal(oracle.collision_counts(h).astype(np.int64), oracle_np.collision_counts(keys))
    for fast in (1, 0):
        idx, dist, nv = oracle.kneighbors(h, 6, fast)
        i2, d2, n2 = oracle_np.kneighbors(Xe, keys, 6, fast=bool(fast), min_blocks=3)
        assert np.array_equal(nv, n2) and np.array_equal(idx, i2)
        fin = np.isfinite(d2)
        np.testing.assert_allclose(dist[fin], d2[fin], rtol=1e-6)
    with pytest.raises(SchicError):
        oracle.set_active_hashes(h, 100)
    ref = oracle.kneighbors(h, 6, 0)
    oracle.destroy(h)
    h = oracle.create(n_neighbors=20, fast=False, **kw)
    for a, b, app in ((0, 8, False), (8, 15, True), (15, 21, True)):
        P = Xe[a:b]
        oracle.fit_csr(h, P.indptr, P.indices, P.data.astype(np.float32), app)
    assert np.array_equal(oracle.signatures(h, T), keys)
    got = oracle.kneighbors(h, 6, 0)
    assert all(np.array_equal(x, y) for x, y in zip(got, ref))
    oracle.destroy(h)
    # the combine changes every key, and the fixture's similarities no longer reach 100 of 160 tables (SURVEY's argument
    # for "off"): far fewer pairs pass the reference's threshold than with raw minima
    cnt = oracle_np.collision_counts(keys)
    assert (cnt[np.triu_indices(21, 1)] >= 100).sum() == 0
