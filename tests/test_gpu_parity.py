"""GPU suite (-m gpu): the sm_100a kernels, called through the C ABI, against the CPU oracle on
the same inputs. Bar: bit-exact for signatures, collision counts and neighbour indices;
relative 1e-5 for reranked distances (BASELINE.json north_star)."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix, vstack

from conftest import REF_KWARGS, Fitted
from schicexplorer_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-5   # north_star: "reranked distances within 1e-5 relative"
# the two counts kernels of k4c_rerank.cu (direct: default; batch iterator: SCHIC_K4_KERNEL=iter or SCHIC_K4_LPS=0|2)
COUNTS_KERNELS = ("k4c_rerank_direct_kernel", "k4c_rerank_kernel")


def synth_csr(n, seed=11, **kw):
    cfg = synth.small_config(n, config_id=seed, **kw)
    ip, ix, dt, z = synth.generate_csr(cfg, workers=1)
    return synth.to_scipy(cfg, ip, ix, dt)


def assert_neighbours_equal_up_to_ties(idx, dist, nv, ridx, rdist, rnv, rtol):
    assert np.array_equal(nv, rnv)
    for c in range(len(nv)):
        k = nv[c]
        np.testing.assert_allclose(dist[c, :k], rdist[c, :k], rtol=rtol, atol=0)
        if not np.array_equal(idx[c, :k], ridx[c, :k]):
            # allowed only where distances tie (within tolerance): same sets per tie group
            assert sorted(idx[c, :k].tolist()) == sorted(ridx[c, :k].tolist())
            bad = idx[c, :k] != ridx[c, :k]
            d = rdist[c, :k][bad]
            assert np.allclose(dist[c, :k][bad], d, rtol=rtol)
        assert (idx[c, k:] == -1).all() and np.isinf(dist[c, k:]).all()


@pytest.mark.parametrize("width", [32, 64])
def test_fixture_signatures_counts_bit_exact(cuda, oracle, fixture20, golden20, width):
    X = fixture20[0]
    with Fitted(cuda, X, hash_width=width) as g:
        sig = g.sig()
        assert np.array_equal(sig, golden20[f"sig{width}"])
        assert np.array_equal(cuda.collision_counts(g.h), golden20[f"cnt{width}"])
        assert np.array_equal(cuda.collision_counts(g.h, 3, 11), golden20[f"cnt{width}"][3:11])


@pytest.mark.parametrize("fast", [True, False])
def test_fixture_kneighbors_and_graph(cuda, oracle, fixture20, golden20, fast):
    X = fixture20[0]
    tag = "fast" if fast else "exact"
    with Fitted(cuda, X, fast=fast) as g, Fitted(oracle, X, fast=fast) as o:
        idx, dist, nv = cuda.kneighbors(g.h, 20)
        assert_neighbours_equal_up_to_ties(idx, dist, nv, golden20[f"idx_{tag}"], golden20[f"dist_{tag}"],
                                           golden20[f"nv_{tag}"], 0 if fast else RTOL)
        if fast:
            assert np.array_equal(idx, golden20["idx_fast"]) and np.array_equal(dist, golden20["dist_fast"])
        gi, gc, gd = cuda.kneighbors_graph(g.h, 20)
        oi, oc, od = oracle.kneighbors_graph(o.h, 20)
        assert np.array_equal(gi, oi) and np.array_equal(gc, oc)
        np.testing.assert_allclose(gd, od, rtol=0 if fast else RTOL)
        # k smaller than / larger than n
        i5, d5, n5 = cuda.kneighbors(g.h, 5)
        oi5, od5, on5 = oracle.kneighbors(o.h, 5)
        assert_neighbours_equal_up_to_ties(i5, d5, n5, oi5, od5, on5, 0 if fast else RTOL)
        i30, d30, n30 = cuda.kneighbors(g.h, 30)
        assert np.array_equal(n30, nv) and (i30[:, 20:] == -1).all()
        assert np.array_equal(i30[:, :20], idx)


@pytest.mark.parametrize("H", [1, 31, 32, 33, 160, 200, 800, 1200, 2000])
def test_signatures_hash_count_sweep(cuda, oracle, H):
    X = synth_csr(48, median_nnz=2500, min_nnz=1, max_nnz=9000)
    with Fitted(cuda, X, number_of_hash_functions=H, minimal_blocks_in_common=1) as g, \
            Fitted(oracle, X, number_of_hash_functions=H, minimal_blocks_in_common=1) as o:
        assert np.array_equal(g.sig(), o.sig())
        assert np.array_equal(cuda.collision_counts(g.h), oracle.collision_counts(o.h))


def test_ragged_empty_and_tiny_rows(cuda, oracle):
    X = synth_csr(40, median_nnz=3000, min_nnz=1, max_nnz=9000)
    empty = csr_matrix((1, X.shape[1]))
    one = csr_matrix(([3.0], ([0], [12345])), shape=(1, X.shape[1]))
    last = csr_matrix(([1.0, 2.0], ([0, 0], [0, X.shape[1] - 1])), shape=(1, X.shape[1]))
    Xe = vstack([empty, X[:9], empty, empty, one, X[9:], last, empty]).tocsr()
    for fast in (True, False):
        kw = dict(number_of_hash_functions=256, minimal_blocks_in_common=8, fast=fast)
        with Fitted(cuda, Xe, **kw) as g, Fitted(oracle, Xe, **kw) as o:
            sig = g.sig()
            assert np.array_equal(sig, o.sig())
            assert (sig[0] == 2147483647).all()
            assert np.array_equal(cuda.collision_counts(g.h), oracle.collision_counts(o.h))
            a, b = cuda.kneighbors(g.h, 6), oracle.kneighbors(o.h, 6)
            assert_neighbours_equal_up_to_ties(*a, *b, 0 if fast else RTOL)


def test_modulo_special_values_hit_the_exact_path(cuda, oracle):
    """Features engineered so that wang32 output is 2M, 2M+1 or M (residues 0, 1, 0): the rare
    outcomes of the two-minima trick in K1 must still be bit-exact."""
    M = 2147483647

    def inv_wang32(k):   # exact inverse of the mix, to find keys whose hash output is a chosen value
        k ^= k >> 16
        k = (k * pow(2057, -1, 1 << 32)) & 0xFFFFFFFF
        k ^= k >> 4; k ^= k >> 8; k ^= k >> 16
        k = (k * pow(5, -1, 1 << 32)) & 0xFFFFFFFF
        k ^= k >> 12; k ^= k >> 24
        k = ((k + 1) * pow(32767, -1, 1 << 32)) & 0xFFFFFFFF
        return k
    feats = []
    for target in (2 * M, 2 * M + 1, M, M - 1, M + 1, 0, 1):
        key = inv_wang32(target)               # (f+1)*(j+1) == key for j = 0  ->  f = key - 1
        f = (key - 1) & 0xFFFFFFFF
        if f < 0xFFFFFFFF:
            feats.append(f)
    rows = []
    ncols = 2 ** 32
    for f in feats:
        rows.append(csr_matrix(([1.0, 1.0], ([0, 0], sorted({f, 777}))), shape=(1, ncols)) if f != 777
                    else csr_matrix(([1.0], ([0], [777])), shape=(1, ncols)))
    X = vstack(rows).tocsr()
    X.indices = X.indices.astype(np.int64)
    with Fitted(cuda, X, number_of_hash_functions=64, minimal_blocks_in_common=1) as g, \
            Fitted(oracle, X, number_of_hash_functions=64, minimal_blocks_in_common=1) as o:
        sg, so = g.sig(), o.sig()
        assert np.array_equal(sg, so)
        assert sorted(so[:3, 0].tolist()) == [0, 0, 1]     # the engineered residues really occur
        assert np.array_equal(cuda.collision_counts(g.h), oracle.collision_counts(o.h))


def test_synthetic_clusters_fast_and_exact(cuda, oracle):
    X = synth_csr(300, seed=5)
    for fast in (True, False):
        kw = dict(fast=fast, n_neighbors=25)
        with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
            assert np.array_equal(g.sig(), o.sig())
            assert np.array_equal(cuda.collision_counts(g.h), oracle.collision_counts(o.h))
            a, b = cuda.kneighbors(g.h, 25), oracle.kneighbors(o.h, 25)
            assert_neighbours_equal_up_to_ties(*a, *b, 0 if fast else RTOL)
            if fast:
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_excess_factor_and_absolute_numbers(cuda, oracle):
    X = synth_csr(120, seed=6)
    kw = dict(fast=False, excess_factor=3, minimal_blocks_in_common=40, n_neighbors=10)
    with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
        a, b = cuda.kneighbors(g.h, 10), oracle.kneighbors(o.h, 10)
        assert_neighbours_equal_up_to_ties(*a, *b, RTOL)
    kw = dict(fast=True, absolute_numbers=True, minimal_blocks_in_common=40)
    with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
        a, b = cuda.kneighbors(g.h, 10), oracle.kneighbors(o.h, 10)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_partial_fit_equals_fit(cuda, fixture20, golden20):
    X = fixture20[0]
    h = cuda.create(n_neighbors=20, fast=False, **REF_KWARGS)
    for a, b, app in ((0, 7, False), (7, 15, True), (15, 20, True)):
        cuda.fit_csr(h, X[a:b].indptr, X[a:b].indices, X[a:b].data.astype(np.float32), app)
    assert cuda.n_fitted(h) == 20
    assert np.array_equal(cuda.signatures(h, 800), golden20["sig32"])
    idx, dist, nv = cuda.kneighbors(h, 20)
    assert_neighbours_equal_up_to_ties(idx, dist, nv, golden20["idx_exact"], golden20["dist_exact"], golden20["nv_exact"], RTOL)
    # refit replaces
    cuda.fit_csr(h, X[:5].indptr, X[:5].indices, X[:5].data.astype(np.float32), False)
    assert cuda.n_fitted(h) == 5 and np.array_equal(cuda.signatures(h, 800), golden20["sig32"][:5])
    cuda.destroy(h)


def test_minhash_class_runs_on_the_gpu(cuda, fixture20, golden20):
    """The reference-facing Python class, unpatched: must be served by the CUDA library."""
    from schicexplorer_b200 import MinHash
    X = fixture20[0]
    mh = MinHash(n_neighbors=20, number_of_hash_functions=800, number_of_cores=8, shingle_size=5, fast=True,
                 maxFeatures=int(max(X.getnnz(1))), absolute_numbers=False, max_bin_size=100000,
                 minimal_blocks_in_common=100, excess_factor=1, prune_inverse_index=False)
    mh.fit(X)
    assert mh._api.backend == "cuda"
    assert np.array_equal(mh.signatures(), golden20["sig32"])
    g = mh.kneighbors_graph(mode="distance")
    assert g.shape == (20, 20) and np.array_equal(np.diff(g.indptr), golden20["nv_fast"])
    assert mh._api.launch_count(mh._h) > 0
    mh.close()


@pytest.mark.parametrize("width", [32, 64])
def test_block_combine_shingle(cuda, oracle, fixture20, width):
    """K2 (k2_combine.cu), the optional block combine behind schic_mh_params.shingle (SURVEY 8(c) rule 4, off by default):
    keys, counts, fast and exact neighbours equal the oracle's for both HashSpecs; empty rows keep the sentinel;
    partial_fit == fit; the device-resident fit entry point; the reference-facing class; the prefix switch is refused."""
    import torch
    from schicexplorer_b200 import MinHash
    from schicexplorer_b200._capi import SchicError
    X = fixture20[0]
    Xe = vstack([X[:9], csr_matrix((1, X.shape[1])), X[9:]]).tocsr()
    H, s = 803, 5
    T = H // s
    kw = dict(number_of_hash_functions=H, shingle=1, shingle_size=s, minimal_blocks_in_common=3, hash_width=width)
    o = oracle.create(n_neighbors=20, fast=False, **kw)
    oracle.fit_csr(o, Xe.indptr, Xe.indices, Xe.data.astype(np.float32), False)
    keys, counts = oracle.signatures(o, T), oracle.collision_counts(o)
    assert (keys[9] == 2147483647).all()
    g = cuda.create(n_neighbors=20, fast=False, **kw)
    cuda.fit_csr(g, Xe.indptr, Xe.indices, Xe.data.astype(np.float32), False)
    assert np.array_equal(cuda.signatures(g, T), keys)
    assert np.array_equal(cuda.collision_counts(g), counts)
    for fast in (1, 0):
        assert_neighbours_equal_up_to_ties(*cuda.kneighbors(g, 6, fast), *oracle.kneighbors(o, 6, fast), RTOL)
    with pytest.raises(SchicError):
        cuda.set_active_hashes(g, 100)
    # a second fit on the same handle (raw matrix restored, combined again), then partial fits
    cuda.fit_csr(g, Xe.indptr, Xe.indices, Xe.data.astype(np.float32), False)
    assert np.array_equal(cuda.signatures(g, T), keys)
    for a, b, app in ((0, 8, False), (8, 15, True), (15, 21, True)):
        P = Xe[a:b]
        cuda.fit_csr(g, P.indptr, P.indices, P.data.astype(np.float32), app)
    assert cuda.n_fitted(g) == 21 and np.array_equal(cuda.signatures(g, T), keys)
    assert_neighbours_equal_up_to_ties(*cuda.kneighbors(g, 6, 0), *oracle.kneighbors(o, 6, 0), RTOL)
    cuda.destroy(g)
    # device-resident CSR
    dev = torch.device("cuda:0")
    ip = torch.from_numpy(Xe.indptr.astype(np.int64)).to(dev)
    ix = torch.from_numpy(Xe.indices.astype(np.int32)).to(dev)
    dt = torch.from_numpy(Xe.data.astype(np.float32)).to(dev)
    g = cuda.create(n_neighbors=20, fast=False, device=0, **kw)
    cuda.fit_csr_device(g, Xe.shape[0], Xe.nnz, ip.data_ptr(), ix.data_ptr(), dt.data_ptr(), False)
    assert np.array_equal(cuda.signatures(g, T), keys)
    cuda.destroy(g)
    oracle.destroy(o)
    # the class: signatures() has one column per table
    mh = MinHash(n_neighbors=20, fast=True, **kw)
    mh.fit(Xe)
    assert mh.signatures().shape == (21, T) and np.array_equal(mh.signatures(), keys)
    with pytest.raises(ValueError):
        mh.set_number_of_hash_functions(100)
    mh.close()


def test_device_resident_entry_points(cuda, oracle):
    import torch
    X = synth_csr(200, seed=8)
    dev = torch.device("cuda:0")
    ip = torch.from_numpy(X.indptr.astype(np.int64)).to(dev)
    ix = torch.from_numpy(X.indices.astype(np.int32)).to(dev)
    dt = torch.from_numpy(X.data.astype(np.float32)).to(dev)
    k = 16
    with Fitted(oracle, X, fast=False, n_neighbors=k) as o:
        h = cuda.create(n_neighbors=k, fast=False, device=0, **REF_KWARGS)
        cuda.set_stream(h, torch.cuda.current_stream().cuda_stream)
        cuda.fit_csr_device(h, X.shape[0], X.nnz, ip.data_ptr(), ix.data_ptr(), dt.data_ptr(), False)
        idx = torch.empty((200, k), dtype=torch.int32, device=dev)
        dist = torch.empty((200, k), dtype=torch.float32, device=dev)
        nv = torch.empty(200, dtype=torch.int32, device=dev)
        cuda.kneighbors_device(h, k, -1, idx.data_ptr(), dist.data_ptr(), nv.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(cuda.signatures(h, 800), o.sig())
        b = oracle.kneighbors(o.h, k)
        assert_neighbours_equal_up_to_ties(idx.cpu().numpy(), dist.cpu().numpy(), nv.cpu().numpy(), *b, RTOL)
        ms = cuda.stage_ms(h)
        assert ms["signatures"] > 0 and ms["collisions"] > 0 and ms["rerank"] > 0
        cuda.destroy(h)


def test_external_database_matches_local(cuda, oracle):
    """Sharded query path on one GPU: two handles each hash half of the rows, the 'gathered'
    signature matrix / CSR is handed back, each answers for its own rows."""
    import torch
    X = synth_csr(150, seed=9)
    dev = torch.device("cuda:0")
    k = 12
    ip = torch.from_numpy(X.indptr.astype(np.int64)).to(dev)
    ix = torch.from_numpy(X.indices.astype(np.int32)).to(dev)
    dt = torch.from_numpy(X.data.astype(np.float32)).to(dev)
    with Fitted(oracle, X, fast=False, n_neighbors=k) as o:
        ref = oracle.kneighbors(o.h, k)
    sig_all = torch.empty((150, 800), dtype=torch.int32, device=dev)
    parts, handles = [(0, 70), (70, 150)], []
    for a, b in parts:
        h = cuda.create(n_neighbors=k, fast=False, device=0, **REF_KWARGS)
        S = X[a:b]
        cuda.fit_csr(h, S.indptr, S.indices, S.data.astype(np.float32), False)
        p, n = cuda.signatures_device(h)
        assert n == b - a
        cuda.synchronize(h)
        src = torch.from_numpy(cuda.signatures(h, 800).view(np.int32)).to(dev)
        sig_all[a:b] = src
        handles.append(h)
    torch.cuda.synchronize()
    for (a, b), h in zip(parts, handles):
        cuda.set_database_device(h, 150, a, sig_all.data_ptr(), ip.data_ptr(), ix.data_ptr(), dt.data_ptr())
        idx, dist, nv = cuda.kneighbors(h, k, n_rows=b - a)
        assert_neighbours_equal_up_to_ties(idx, dist, nv, ref[0][a:b], ref[1][a:b], ref[2][a:b], RTOL)
        cuda.destroy(h)


def test_full_size_properties(cuda):
    """Size-independent properties at BASELINE scale per cell (full-length rows, H=800): order
    invariance, partial_fit == fit, idempotence, symmetry of counts, self collision = H."""
    cfg = synth.SynthConfig(3, 256)
    ip, ix, dt, _ = synth.generate_csr(cfg, 0, 256, workers=1)
    X = synth.to_scipy(cfg, ip, ix, dt)
    with Fitted(cuda, X, fast=True, n_neighbors=50) as g:
        sig = g.sig()
        cnt = cuda.collision_counts(g.h)
        assert np.array_equal(cnt, cnt.T) and (np.diag(cnt) == 800).all()
        assert sig.max() < 2147483647
        # idempotence
        cuda.fit_csr(g.h, X.indptr, X.indices, None, False)
        assert np.array_equal(g.sig(), sig)
        # permutation of the non-zeros inside rows
        rng = np.random.default_rng(3)
        perm = X.indices.copy()
        for r in range(0, 256, 17):
            a, b = X.indptr[r], X.indptr[r + 1]
            perm[a:b] = rng.permutation(perm[a:b])
        cuda.fit_csr(g.h, X.indptr, perm, None, False)
        assert np.array_equal(g.sig(), sig)
        # fit in two appends
        cuda.fit_csr(g.h, X[:100].indptr, X[:100].indices, None, False)
        cuda.fit_csr(g.h, X[100:].indptr, X[100:].indices, None, True)
        assert np.array_equal(g.sig(), sig)
        # signature of a row is the column-wise min of the signatures of any split of its features
        r = 5
        a, b = X.indptr[r], X.indptr[r + 1]
        mid = (a + b) // 2
        halves = csr_matrix((np.ones(b - a), X.indices[a:b], [0, mid - a, b - a]), shape=(2, X.shape[1]))
        cuda.fit_csr(g.h, halves.indptr, halves.indices, None, False)
        assert np.array_equal(g.sig().min(axis=0), sig[r])


def test_error_codes(cuda):
    from schicexplorer_b200._capi import SchicError
    with pytest.raises(SchicError):
        cuda.create(n_neighbors=5, number_of_hash_functions=70000)
    h = cuda.create(n_neighbors=5)
    with pytest.raises(SchicError) as e:
        cuda.kneighbors(h, 5, n_rows=1)
    assert e.value.status == -2
    cuda.destroy(h)


def test_multi_gpu_shard_and_gather_is_identical():
    """NCCL path: needs >= 2 GPUs on the box (skipped on a 1-GPU box; tools/verify_multi_gpu.py is the
    same check run by hand under torchrun, and tests/test_distributed.py covers the host logic on gloo)."""
    import os
    import subprocess
    import sys
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.join(os.path.dirname(__file__), "..")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577",
                        os.path.join(root, "tools", "verify_multi_gpu.py")], capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_PARITY_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_max_bin_size_bucket_pruning(cuda, oracle):
    """SURVEY 8(c) rule 5: buckets with >= max_bin_size members are ignored (K2)."""
    X = synth_csr(1, median_nnz=300, min_nnz=100, max_nnz=600)
    Xr = vstack([X] * 6).tocsr()
    for mbs, expect in ((5, 0), (6, 0), (7, 32)):
        with Fitted(cuda, Xr, number_of_hash_functions=32, minimal_blocks_in_common=1, max_bin_size=mbs) as g:
            assert (cuda.collision_counts(g.h) == expect).all()
    # the --saveMemory branch of the reference does not pass max_bin_size (upstream default 50):
    # with a few hundred clustered cells many buckets are larger than that
    Xs = synth_csr(260, seed=13)
    for fast in (True, False):
        kw = dict(fast=fast, n_neighbors=15, max_bin_size=50, minimal_blocks_in_common=20)
        with Fitted(cuda, Xs, **kw) as g, Fitted(oracle, Xs, **kw) as o:
            cg, co = cuda.collision_counts(g.h), oracle.collision_counts(o.h)
            assert np.array_equal(cg, co)
            assert (np.diag(co) < 800).any()          # pruning really happened
            a, b = cuda.kneighbors(g.h, 15), oracle.kneighbors(o.h, 15)
            assert_neighbours_equal_up_to_ties(*a, *b, 0 if fast else RTOL)


@pytest.mark.parametrize("mode", ["counts", "hash", "window", "window64"])
def test_rerank_kernel_variants(cuda, oracle, monkeypatch, mode):
    """K4 has three implementations (counts kernel: packed stream + byte table, the default for integer counts; hash
    table; windowed direct table with 32- or 64-bit entries); SCHIC_K4_MODE selects one ("counts" = default: counts
    kernel with the generic ones behind it). All must agree with the oracle, including on rows that
    (a) hold more than one table fill (> 4096 non-zeros) inside ONE window, (b) are empty / tiny,
    (c) use ids near 2^32 (too many windows: the library must fall back to the hash kernel)."""
    monkeypatch.setenv("SCHIC_K4_MODE", mode)
    rng = np.random.default_rng(17)
    # (a)+(b): 60 rows drawn from a 2^18-wide id range starting mid-window, 5000..12000 non-zeros each
    ncols = 1 << 23
    base = (5 << 18) + 1000
    rows, cols, vals = [], [], []
    pool = rng.choice(200000, size=40000, replace=False) + base
    for r in range(60):
        m = int(rng.integers(5000, 12000))
        c = np.sort(rng.choice(pool, size=m, replace=False))
        extra = np.sort(rng.choice(ncols, size=300, replace=False))      # a few ids in other windows
        c = np.unique(np.concatenate([c, extra]))
        rows.append(np.full(len(c), r)); cols.append(c); vals.append(rng.integers(1, 50, size=len(c)).astype(np.float64))
    X = csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(60, ncols))
    empty = csr_matrix((1, ncols))
    one = csr_matrix(([3.0], ([0], [base + 7])), shape=(1, ncols))
    X = vstack([empty, X[:30], one, X[30:], empty]).tocsr()
    kw = dict(fast=False, n_neighbors=20, number_of_hash_functions=256, minimal_blocks_in_common=1)
    with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
        a, b = cuda.kneighbors(g.h, 20), oracle.kneighbors(o.h, 20)
        assert_neighbours_equal_up_to_ties(*a, *b, RTOL)
        assert cuda.rerank_kernel_name(g.h) in {"counts": COUNTS_KERNELS, "hash": ("k4_rerank_kernel",)}.get(mode, ("k4_rerank_window_kernel",))
    # realistic clustered cells, k*excess_factor candidates
    Xs = synth_csr(200, seed=21)
    kw = dict(fast=False, n_neighbors=15, excess_factor=2, minimal_blocks_in_common=30)
    with Fitted(cuda, Xs, **kw) as g, Fitted(oracle, Xs, **kw) as o:
        a, b = cuda.kneighbors(g.h, 15), oracle.kneighbors(o.h, 15)
        assert_neighbours_equal_up_to_ties(*a, *b, RTOL)
    # (c) ids spread over the whole 32-bit range
    ncols = 2 ** 32
    rows, cols = [], []
    shared = np.sort(rng.choice(ncols - 1, size=3000, replace=False))
    for r in range(24):
        c = np.unique(np.concatenate([rng.choice(shared, size=2000, replace=False),
                                      rng.choice(ncols - 1, size=500, replace=False), [ncols - 1]]))
        rows.append(np.full(len(c), r)); cols.append(c)
    cc = np.concatenate(cols)
    Xw = csr_matrix((rng.integers(1, 9, size=len(cc)).astype(np.float64), (np.concatenate(rows), cc)), shape=(24, ncols))
    Xw.indices = Xw.indices.astype(np.int64)
    kw = dict(fast=False, n_neighbors=8, number_of_hash_functions=128, minimal_blocks_in_common=1)
    with Fitted(cuda, Xw, **kw) as g, Fitted(oracle, Xw, **kw) as o:
        a, b = cuda.kneighbors(g.h, 8), oracle.kneighbors(o.h, 8)
        assert_neighbours_equal_up_to_ties(*a, *b, RTOL)


def _count_rows(rng, n_rows, ncols, lens, hi=40, pool=None):
    rows, cols, vals = [], [], []
    for r in range(n_rows):
        m = int(lens[r % len(lens)])
        src = ncols if pool is None else pool
        c = np.sort(rng.choice(src, size=min(m, src if np.isscalar(src) else len(src)), replace=False))
        rows.append(np.full(len(c), r)); cols.append(c); vals.append(rng.integers(1, hi, size=len(c)).astype(np.float64))
    return csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n_rows, ncols))


def Xa_early(X, rng):
    """X with a few counts >= 255 (escape windows) and slices that start / end on every alignment."""
    Xa = X.copy()
    big = rng.choice(Xa.nnz, size=Xa.nnz // 300, replace=False)
    Xa.data[big] = rng.integers(255, 65536, size=len(big)).astype(np.float64)
    return Xa


def test_counts_rerank_kernel(cuda, oracle, monkeypatch):
    """The counts kernel (k4c_rerank.cu) on its own edge cases, always against the oracle:
    (a) counts >= 255 in query rows (table bytes saturate at 255, resolved by the escape search), up to the largest
        packable count 65535; (b) windows with more than 2048 query keys (build / un-build loop past the two
        prefetched keys); (c) window groups: one group, and one window per group (many CTAs per query, global
        accumulation); (d) data the kernel must hand to the generic float kernels -- non-integer values, a count above
        65535, more than 1/64 of the values >= 255 -- decided by the device flag, with identical results;
    (e) brute-force exact kNN (every row a candidate, 1024-row chunks) through the same kernel."""
    rng = np.random.default_rng(41)
    ncols = 1 << 22
    # clustered rows: a shared pool inside ONE 2^17 window (rows with 3000-6000 keys there) plus scattered ids
    pool = np.sort(rng.choice(1 << 17, size=20000, replace=False)) + (3 << 17)
    base = _count_rows(rng, 70, ncols, [3000, 4500, 6000, 800, 2049, 2048], pool=pool)
    scat = _count_rows(rng, 70, ncols, [700, 50, 1200])
    X = (base + scat).tocsr()
    X.sum_duplicates()
    X.sort_indices()
    kw = dict(fast=False, n_neighbors=25, number_of_hash_functions=128, minimal_blocks_in_common=1, excess_factor=2)

    def check(M, want_kernel, k=25):
        with Fitted(cuda, M, **kw) as g, Fitted(oracle, M, **kw) as o:
            cuda.set_feature_range(g.h, M.shape[1])
            a, b = cuda.kneighbors(g.h, k), oracle.kneighbors(o.h, k)
            assert_neighbours_equal_up_to_ties(*a, *b, RTOL)
            assert cuda.rerank_kernel_name(g.h) in ((want_kernel,) if isinstance(want_kernel, str) else want_kernel)
            return a

    for wg in ("1", "3", "100000", None):                                # (c): one window per group ... one group / default
        if wg is None:
            monkeypatch.delenv("SCHIC_K4_GROUP_WINDOWS", raising=False)
        else:
            monkeypatch.setenv("SCHIC_K4_GROUP_WINDOWS", wg)
        check(X, COUNTS_KERNELS)
    for lps in ("8", "4", "0", "2"):                                     # lane-group variants (0: proportional to row length)
        monkeypatch.setenv("SCHIC_K4_LPS", lps)
        for kern in ("direct", "iter"):                                  # (direct exists for groups of 4 and 8 lanes)
            monkeypatch.setenv("SCHIC_K4_KERNEL", kern)
            want = "k4c_rerank_direct_kernel" if kern == "direct" and lps in ("8", "4") else "k4c_rerank_kernel"
            check(X, want)
            check(X[:9], want, k=9)
    monkeypatch.delenv("SCHIC_K4_LPS")
    # direct kernel: every (lanes, chunks per round, threads) build
    for lps, u in (("8", "2"), ("8", "3"), ("4", "4"), ("4", "6")):
        for threads in ("416", "512"):
            monkeypatch.setenv("SCHIC_K4_KERNEL", "direct")
            monkeypatch.setenv("SCHIC_K4_U", u)
            monkeypatch.setenv("SCHIC_K4_LPS", lps)
            monkeypatch.setenv("SCHIC_K4_THREADS", threads)
            check(X, "k4c_rerank_direct_kernel")
            check(Xa_early(X, rng), "k4c_rerank_direct_kernel")
    for name in ("SCHIC_K4_U", "SCHIC_K4_LPS", "SCHIC_K4_KERNEL", "SCHIC_K4_THREADS"):
        monkeypatch.delenv(name)
    # (a) a few large counts (escapes) in some rows, incl. the largest packable value
    Xa = X.copy()
    big = rng.choice(Xa.nnz, size=Xa.nnz // 200, replace=False)
    Xa.data[big] = rng.integers(255, 65536, size=len(big)).astype(np.float64)
    Xa.data[big[:3]] = [255.0, 256.0, 65535.0]
    check(Xa, COUNTS_KERNELS)
    # (d) not packable: every variant must come out of the generic kernel, same answers as the oracle
    Xf = X.copy(); Xf.data = Xf.data * 0.5                               # non-integer values
    check(Xf, "k4_rerank_window_kernel")
    Xb = X.copy(); Xb.data[7] = 65536.0                                  # one count beyond 16 bits
    check(Xb, "k4_rerank_window_kernel")
    Xm = X.copy(); Xm.data[rng.choice(Xm.nnz, size=Xm.nnz // 20, replace=False)] = 300.0   # too many escapes
    check(Xm, "k4_rerank_window_kernel")
    Xn = X.copy(); Xn.data[11] = -1.0                                    # negative
    check(Xn, "k4_rerank_window_kernel")
    # (e) exact kNN (brute force) through the counts kernel == oracle == the generic kernels
    with Fitted(cuda, Xa, **kw) as g, Fitted(oracle, Xa, **kw) as o:
        gi, gd, gn = cuda.kneighbors_exact(g.h, 10, False)
        oi, od, on = oracle.kneighbors_exact(o.h, 10, False)
        assert cuda.rerank_kernel_name(g.h) in COUNTS_KERNELS
        assert_neighbours_equal_up_to_ties(gi, gd, gn, oi, od, on, RTOL)
        monkeypatch.setenv("SCHIC_K4_MODE", "window")
    with Fitted(cuda, Xa, **kw) as g:
        wi, wd, wn = cuda.kneighbors_exact(g.h, 10, False)
        assert cuda.rerank_kernel_name(g.h) == "k4_rerank_window_kernel"
        assert_neighbours_equal_up_to_ties(gi, gd, gn, wi, wd, wn, RTOL)


def test_k1_filtered_equals_chunked_kernel(cuda, oracle, monkeypatch):
    """K1 has two implementations: the filtered whole-row kernel (default) and the unfiltered fixed-chunk
    kernel (SCHIC_K1_MODE=chunk). Both must be bit-identical to the oracle on rows of every length class:
    empty, 1..13 non-zeros (no speculation), a few hundred, > 16384 (split into parts), > 2048 (several tiles)."""
    rng = np.random.default_rng(23)
    ncols = 2744 * 2744
    lens = [0, 1, 2, 3, 4, 5, 11, 12, 13, 14, 63, 64, 65, 500, 2047, 2048, 2049, 4100, 16383, 16384, 16385, 40000, 0, 70001]
    rows, cols = [], []
    for r, m in enumerate(lens):
        c = np.sort(rng.choice(ncols, size=m, replace=False))
        rows.append(np.full(m, r)); cols.append(c)
    cc = np.concatenate(cols)
    X = csr_matrix((np.ones(len(cc)), (np.concatenate(rows), cc)), shape=(len(lens), ncols))
    for H in (800, 1200, 96):
        kw = dict(number_of_hash_functions=H, minimal_blocks_in_common=1)
        with Fitted(oracle, X, **kw) as o:
            want = o.sig()
        for mode in ("tile", "direct", "chunk", "table"):
            monkeypatch.setenv("SCHIC_K1_MODE", mode)
            with Fitted(cuda, X, **kw) as g:
                assert np.array_equal(g.sig(), want), (H, mode)


@pytest.mark.parametrize("width", [32, 64])
@pytest.mark.parametrize("H", [800, 4000, 33])
def test_k1_shared_table_path(cuda, oracle, monkeypatch, H, width):
    """The shared-table K1 (every (id, hash function) evaluated once per batch, per-cell lookup, exact fallback)
    on clustered cells that really share ids, forced on (the automatic choice needs >= 2M non-zeros): signatures
    and collision counts bit-identical to the oracle; also through partial_fit (a table per appended batch)."""
    monkeypatch.setenv("SCHIC_K1_MODE", "table")
    X = synth_csr(160, seed=29)
    empty = csr_matrix((1, X.shape[1]))
    Xe = vstack([X[:50], empty, X[50:]]).tocsr()
    kw = dict(number_of_hash_functions=H, minimal_blocks_in_common=1, hash_width=width)
    with Fitted(oracle, Xe, **kw) as o:
        want, want_cnt = o.sig(), oracle.collision_counts(o.h)
    with Fitted(cuda, Xe, **kw) as g:
        assert np.array_equal(g.sig(), want)
        assert np.array_equal(cuda.collision_counts(g.h), want_cnt)
        cuda.fit_csr(g.h, Xe[:70].indptr, Xe[:70].indices, None, False)
        cuda.fit_csr(g.h, Xe[70:].indptr, Xe[70:].indices, None, True)
        assert np.array_equal(g.sig(), want)


def test_k1_whole_range_table_is_kept_across_fits(cuda, oracle, monkeypatch):
    """With the feature-range hint the shared K1 table can be built for EVERY id of the range; it then depends only on
    (range, H, hash width), so the handle keeps it: first fit of a shape builds from its own ids (state 1), the second
    builds the whole-range table (2), later ones -- also on DIFFERENT data of that shape -- only look up (3). Every fit
    bit-identical to the oracle; SCHIC_K1_NO_CACHE switches the re-use off; another H or range invalidates."""
    from schicexplorer_b200 import synth
    import os
    # full-length S10-like rows; the automatic table path needs >= 2 M non-zeros AND an id range <= nnz / 2
    cfg = synth.SynthConfig(71, 1150)
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=min(16, os.cpu_count() or 1))
    cfg2 = synth.SynthConfig(72, 1100)
    ip2, ix2, dt2, _ = synth.generate_csr(cfg2, workers=min(16, os.cpu_count() or 1))
    assert ip[-1] > 2 * cfg.chroms.nbins ** 2 and ip2[-1] > 2 * cfg.chroms.nbins ** 2
    kw = dict(n_neighbors=5, fast=True, **REF_KWARGS)

    def oracle_sig(a, b):
        ho = oracle.create(number_of_cores=os.cpu_count() or 1, **kw)
        oracle.fit_csr(ho, a, b, None, False)
        sg = oracle.signatures(ho, 800)
        oracle.destroy(ho)
        return sg

    want, want2 = oracle_sig(ip, ix), oracle_sig(ip2, ix2)
    h = cuda.create(**kw)
    cuda.set_feature_range(h, cfg.chroms.nbins ** 2)
    states = []
    for a, b, w in ((ip, ix, want), (ip, ix, want), (ip, ix, want), (ip2, ix2, want2), (ip, ix, want)):
        cuda.fit_csr(h, a, b, None, False)
        states.append(cuda.k1_table_state(h))
        assert np.array_equal(cuda.signatures(h, 800), w)
    # (a host fit of this size pre-builds the whole-range table while the ids upload when that is cheap enough; either
    # way the table must be re-used from the third fit on at the latest)
    assert states[2:] == [3, 3, 3] and states[0] in (1, 2) and states[1] in (2, 3), states
    monkeypatch.setenv("SCHIC_K1_NO_CACHE", "1")
    cuda.fit_csr(h, ip, ix, None, False)
    assert cuda.k1_table_state(h) in (1, 2) and np.array_equal(cuda.signatures(h, 800), want)
    monkeypatch.delenv("SCHIC_K1_NO_CACHE")
    cuda.destroy(h)
    # device-resident fits (the benchmark's loop): 1 -> 2 -> 3
    import torch
    dev = torch.device("cuda:0")
    d_ip, d_ix = torch.from_numpy(ip).to(dev), torch.from_numpy(ix.view(np.int32)).to(dev)
    h = cuda.create(**kw)
    cuda.set_feature_range(h, cfg.chroms.nbins ** 2)
    states = []
    for _ in range(4):
        cuda.fit_csr_device(h, len(ip) - 1, int(ip[-1]), d_ip.data_ptr(), d_ix.data_ptr(), None, False, indptr0=0)
        states.append(cuda.k1_table_state(h))
        assert np.array_equal(cuda.signatures(h, 800), want)
    assert states == [1, 2, 3, 3], states
    cuda.destroy(h)


def test_fit_csr_host_staged_upload(cuda, oracle, monkeypatch):
    """schic_mh_fit_csr_host: scipy's own arrays (float64 / float32 data, int32 / int64 indices and indptr, pageable) are
    staged by worker threads, values narrowed per 2^20-element chunk to uint8 / uint16 / float32. Same answers as the
    oracle for: integer counts (all chunks uint8), counts up to 65535 (uint16 chunks), a matrix whose later chunks hold
    non-integer values (mixed wire kinds inside one fit), float32 input, int64 index arrays, append, 1 and many threads."""
    from schicexplorer_b200 import synth
    cfg = synth.SynthConfig(73, 190)                        # ~2.7 M non-zeros: three staging chunks
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=4)
    assert ip[-1] > (2 << 20)
    rng = np.random.default_rng(9)
    kw = dict(n_neighbors=12, fast=False, **REF_KWARGS)

    def reference(data32):
        ho = oracle.create(**kw)
        oracle.fit_csr(ho, ip, ix, data32, False)
        res = oracle.kneighbors(ho, 12)
        oracle.destroy(ho)
        return res

    variants = {}
    variants["counts"] = dt.astype(np.float64)
    big = dt.astype(np.float64).copy(); big[rng.choice(len(big), 5000, replace=False)] = rng.integers(256, 65536, 5000); big[:3] = [255, 256, 65535]
    variants["u16"] = big
    mixed = dt.astype(np.float64).copy(); mixed[(3 << 19):] *= 0.25          # from the middle of chunk 1 on: fractions
    variants["mixed"] = mixed
    over = dt.astype(np.float64).copy(); over[-5] = 65536.0                   # last chunk needs float32
    variants["over"] = over
    for name, data in variants.items():
        want = reference(data.astype(np.float32))
        for threads, idx_dtype, data_dtype in ((0, np.int32, np.float64), (1, np.int64, np.float64), (3, np.int32, np.float32)):
            if data_dtype == np.float32 and not np.array_equal(data.astype(np.float32).astype(np.float64), data):
                continue
            if threads:
                monkeypatch.setenv("SCHIC_STAGE_THREADS", str(threads))
            else:
                monkeypatch.delenv("SCHIC_STAGE_THREADS", raising=False)
            h = cuda.create(**kw)
            cuda.set_feature_range(h, cfg.chroms.nbins ** 2)
            cuda.fit_csr_host(h, ip.astype(np.int32 if idx_dtype == np.int32 else np.int64), ix.astype(idx_dtype),
                              data.astype(data_dtype), False)
            got = cuda.kneighbors(h, 12)
            assert_neighbours_equal_up_to_ties(*got, *want, RTOL)
            assert cuda.rerank_kernel_name(h) in (COUNTS_KERNELS if name in ("counts", "u16") else ("k4_rerank_window_kernel",)), name
            cuda.destroy(h)
    # append (partial_fit) in two staged calls == one fit; fast mode without values
    want = reference(dt)
    h = cuda.create(**kw)
    half = 95
    cuda.fit_csr_host(h, ip[:half + 1], ix, dt.astype(np.float64), False)
    cuda.fit_csr_host(h, ip[half:], ix, dt.astype(np.float64), True)
    assert_neighbours_equal_up_to_ties(*cuda.kneighbors(h, 12), *want, RTOL)
    cuda.destroy(h)
    kwf = dict(kw, fast=True)
    h, ho = cuda.create(**kwf), oracle.create(**kwf)
    cuda.fit_csr_host(h, ip, ix.astype(np.int32), None, False)
    oracle.fit_csr(ho, ip, ix, None, False)
    assert np.array_equal(cuda.signatures(h, 800), oracle.signatures(ho, 800))
    for a, b in zip(cuda.kneighbors(h, 12), oracle.kneighbors(ho, 12)):
        assert np.array_equal(a, b)
    cuda.destroy(h); oracle.destroy(ho)
    # an empty batch
    h = cuda.create(**kwf)
    cuda.fit_csr_host(h, np.zeros(4, dtype=np.int32), np.zeros(0, dtype=np.int32), None, False)
    assert cuda.n_fitted(h) == 3 and (cuda.signatures(h, 800) == 2147483647).all()
    cuda.destroy(h)


def test_collision_blocks_and_given_counts(cuda, oracle):
    """schic_mh_collision_block_device / schic_mh_set_counts_device (the pieces of the sharded symmetric K3): arbitrary
    blocks of the count matrix, with and without the transposed copy, equal the corresponding part of the full matrix
    (= the oracle's); a count matrix assembled from blocks and handed back gives the same neighbours as the built-in K3,
    also with bucket pruning (max_bin_size) active."""
    import torch
    X = synth_csr(300, seed=51)
    n = X.shape[0]
    dev = torch.device("cuda:0")
    for extra in ({}, {"max_bin_size": 40, "minimal_blocks_in_common": 20}):
        kw = dict(fast=True, n_neighbors=15, **extra)
        with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
            cuda.set_stream(g.h, torch.cuda.current_stream().cuda_stream)
            full = oracle.collision_counts(o.h).astype(np.int64)
            want = oracle.kneighbors(o.h, 15)
            for (q0, qn, d0, dn) in ((0, n, 0, n), (17, 100, 130, 77), (130, 77, 17, 100), (5, 1, 0, n), (40, 128, 40, 128), (0, 0, 0, 10)):
                ld, ldt = (dn + 7) // 8 * 8 + 8, (qn + 7) // 8 * 8
                c = torch.full((max(qn, 1), ld), -1, dtype=torch.int16, device=dev)
                ct = torch.full((max(dn, 1), max(ldt, 8)), -1, dtype=torch.int16, device=dev)
                cuda.collision_block_device(g.h, q0, qn, d0, dn, c.data_ptr(), ld, ct.data_ptr(), max(ldt, 8))
                torch.cuda.synchronize()
                got = c.cpu().numpy().view(np.uint16)[:qn, :dn].astype(np.int64)
                assert np.array_equal(got, full[q0:q0 + qn, d0:d0 + dn])
                if not (q0 == d0 and qn == dn) and qn and dn:
                    got_t = ct.cpu().numpy().view(np.uint16)[:dn, :qn].astype(np.int64)
                    assert np.array_equal(got_t, full[q0:q0 + qn, d0:d0 + dn].T)
            # assemble the whole matrix from a 2 x 2 split computed "once per pair" and hand it back
            ld = (n + 7) // 8 * 8
            m = torch.zeros((n, ld), dtype=torch.int16, device=dev)
            h1 = 137
            cuda.collision_block_device(g.h, 0, h1, 0, h1, m.data_ptr(), ld)
            cuda.collision_block_device(g.h, h1, n - h1, h1, n - h1, m.data_ptr() + (h1 * ld + h1) * 2, ld)
            tr = torch.zeros((n - h1, (h1 + 7) // 8 * 8), dtype=torch.int16, device=dev)
            cuda.collision_block_device(g.h, 0, h1, h1, n - h1, m.data_ptr() + h1 * 2, ld, tr.data_ptr(), tr.shape[1])
            m[h1:, :h1].copy_(tr[:, :h1])
            cuda.set_counts_device(g.h, m.data_ptr(), ld)
            got = cuda.kneighbors(g.h, 15)
            for a, b in zip(got, want):
                assert np.array_equal(a, b)
            for a, b in zip(cuda.kneighbors(g.h, 15), want):          # one-shot: the next call computes K3 itself again
                assert np.array_equal(a, b)


def test_long_row_fallbacks(cuda, oracle, monkeypatch):
    """k6_longrows.cu: select, graph emit and rerank merge through the global-memory row sort (what a dense graph of more
    than 16 384 cells takes -- the tool's default k = n), forced on small inputs with SCHIC_FORCE_LONGROWS: identical
    to the oracle in fast and exact mode, k = n and k < n, with excess candidates, chunked rerank lists and the
    brute-force exact kNN; and once for real on rows longer than the shared-memory limit."""
    monkeypatch.setenv("SCHIC_FORCE_LONGROWS", "1")
    X = synth_csr(210, seed=61)
    n = X.shape[0]
    for fast, k, ef in ((True, n, 1), (True, 17, 3), (False, n, 1), (False, 30, 2)):
        kw = dict(fast=fast, n_neighbors=k, excess_factor=ef, minimal_blocks_in_common=40)
        with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
            a, b = cuda.kneighbors(g.h, k), oracle.kneighbors(o.h, k)
            assert_neighbours_equal_up_to_ties(*a, *b, 0 if fast else RTOL)
            gi, gc, gd = cuda.kneighbors_graph(g.h, k)
            oi, oc, od = oracle.kneighbors_graph(o.h, k)
            assert np.array_equal(gi, oi) and np.array_equal(gc, oc)
            np.testing.assert_allclose(gd, od, rtol=0 if fast else RTOL)
    # chunked rerank (lists longer than 1024) + merge, and the brute-force exact kNN, through the global merge
    Xl = synth_csr(1300, seed=62, median_nnz=300, min_nnz=60, max_nnz=900)
    kw = dict(fast=False, n_neighbors=1300, minimal_blocks_in_common=1, number_of_hash_functions=64)
    with Fitted(cuda, Xl, **kw) as g, Fitted(oracle, Xl, **kw) as o:
        assert_neighbours_equal_up_to_ties(*cuda.kneighbors(g.h, 1300), *oracle.kneighbors(o.h, 1300), RTOL)
        a, b = cuda.kneighbors_exact(g.h, 25, False), oracle.kneighbors_exact(o.h, 25, False)
        assert_neighbours_equal_up_to_ties(*a, *b, RTOL)
    monkeypatch.delenv("SCHIC_FORCE_LONGROWS")
    # for real: 17 000 tiny cells, dense graph (k = n > 16 384), fast distances -- graph identical to the oracle's
    rng = np.random.default_rng(63)
    nbig, ncols = 17000, 50000
    pool = rng.choice(ncols, size=400, replace=False)
    rows = np.repeat(np.arange(nbig), 12)
    cols = np.concatenate([np.sort(rng.choice(pool, size=12, replace=False)) for _ in range(nbig)])
    Xb = csr_matrix((np.ones(len(cols)), (rows, cols)), shape=(nbig, ncols))
    kw = dict(fast=True, n_neighbors=nbig, number_of_hash_functions=32, minimal_blocks_in_common=8)
    with Fitted(cuda, Xb, **kw) as g, Fitted(oracle, Xb, **kw) as o:
        gi, gc, gd = cuda.kneighbors_graph(g.h, nbig)
        oi, oc, od = oracle.kneighbors_graph(o.h, nbig)
        assert np.array_equal(gi, oi) and np.array_equal(gc, oc) and np.array_equal(gd, od)
        assert gi[-1] > 5 * nbig


def test_async_host_fit(cuda, oracle):
    """schic_mh_fit_csr_async: returns before the value array has landed; results identical to the blocking call."""
    X = synth_csr(220, seed=33)
    k = 14
    ip = X.indptr.astype(np.int64)
    ix = np.ascontiguousarray(X.indices.astype(np.uint32))
    dt = np.ascontiguousarray(X.data.astype(np.float32))
    with Fitted(oracle, X, fast=False, n_neighbors=k) as o:
        want = oracle.kneighbors(o.h, k)
        want_sig = o.sig()
    h = cuda.create(n_neighbors=k, fast=False, **REF_KWARGS)
    for _ in range(2):                                   # second round: refit while nothing is pending
        cuda.fit_csr(h, ip, ix, dt, False, async_upload=True)
        got = cuda.kneighbors(h, k)
        assert_neighbours_equal_up_to_ties(*got, *want, RTOL)
        assert np.array_equal(cuda.signatures(h, 800), want_sig)
    # async append
    cuda.fit_csr(h, ip[:101], ix[:ip[100]], dt[:ip[100]], False, async_upload=True)
    ip2 = ip[100:] - ip[100]
    cuda.fit_csr(h, ip2, ix[ip[100]:], dt[ip[100]:], True, async_upload=True)
    got = cuda.kneighbors(h, k)
    assert_neighbours_equal_up_to_ties(*got, *want, RTOL)
    with pytest.raises(ValueError):
        cuda.fit_csr(h, X.indptr, X.indices, X.data, False, async_upload=True)   # int32 indptr / float64 data
    cuda.destroy(h)


def test_select_ties_and_overflow_fallback(cuda, oracle):
    """Top-k select: ties at the threshold must keep the lowest indices. 2500 copies of three different rows give
    every query > 2048 candidates with the same count (the single-histogram kernel's tie buffer overflows, the
    radix kernel redoes those rows); n is not a multiple of 8 (row padding of the count matrix)."""
    base = synth_csr(3, median_nnz=600, min_nnz=300, max_nnz=900)
    reps = [base[i % 3] for i in range(7501)]
    X = vstack(reps).tocsr()
    kw = dict(number_of_hash_functions=64, minimal_blocks_in_common=1, n_neighbors=30)
    with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
        a, b = cuda.kneighbors(g.h, 30), oracle.kneighbors(o.h, 30)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # moderate ties (fit the buffer): 40 copies
    X = vstack([base[i % 3] for i in range(121)] + [synth_csr(50, seed=41, median_nnz=600, min_nnz=300, max_nnz=900)]).tocsr()
    with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
        a, b = cuda.kneighbors(g.h, 30), oracle.kneighbors(o.h, 30)
        assert np.array_equal(a[2], b[2]) and np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("H", [800, 801, 4094, 7])
def test_k3_packed_equals_plain(cuda, oracle, monkeypatch, H):
    """K3 has two implementations: 32-bit compares and the packed one (values renamed per hash function to 16-bit
    ids in fp16 patterns, two hash functions per compare). Counts must be identical, including inadmissible
    values (empty rows -> sentinel M), odd H (padding column) and a partial row range."""
    X = synth_csr(300, seed=37)
    empty = csr_matrix((1, X.shape[1]))
    Xe = vstack([X[:100], empty, X[100:], empty]).tocsr()
    kw = dict(number_of_hash_functions=H, minimal_blocks_in_common=1)
    with Fitted(oracle, Xe, **kw) as o:
        want = oracle.collision_counts(o.h)
    got = {}
    for mode in ("plain", "packed"):
        monkeypatch.setenv("SCHIC_K3_MODE", mode)
        with Fitted(cuda, Xe, **kw) as g:
            got[mode] = cuda.collision_counts(g.h)
            assert np.array_equal(cuda.collision_counts(g.h, 37, 203), want[37:203]), mode
        assert np.array_equal(got[mode], want), mode


def test_feature_range_hint(cuda, oracle, monkeypatch):
    """schic_mh_set_feature_range: with the hint nothing is read back during fit / kneighbors and the results are the
    same; an id beyond the declared range is detected on the device and reported (never written out of bounds)."""
    from schicexplorer_b200._capi import SchicError
    monkeypatch.setenv("SCHIC_K1_MODE", "table")
    X = synth_csr(150, seed=43)
    k = 12
    with Fitted(oracle, X, fast=False, n_neighbors=k) as o:
        want, want_sig = oracle.kneighbors(o.h, k), o.sig()
    h = cuda.create(n_neighbors=k, fast=False, **REF_KWARGS)
    cuda.set_feature_range(h, X.shape[1])
    cuda.fit_csr(h, X.indptr, X.indices, X.data.astype(np.float32), False)
    assert np.array_equal(cuda.signatures(h, 800), want_sig)
    assert_neighbours_equal_up_to_ties(*cuda.kneighbors(h, k), *want, RTOL)
    # declared range too small: reported, results of this handle are not to be used
    cuda.set_feature_range(h, int(X.indices.max()) - 5)
    cuda.fit_csr(h, X.indptr, X.indices, X.data.astype(np.float32), False)
    with pytest.raises(SchicError):
        cuda.signatures(h, 800)
    # cleared again: back to the measured range
    cuda.set_feature_range(h, 0)
    cuda.fit_csr(h, X.indptr, X.indices, X.data.astype(np.float32), False)
    assert np.array_equal(cuda.signatures(h, 800), want_sig)
    cuda.destroy(h)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
def test_narrow_count_upload(cuda, oracle, dtype):
    """schic_mh_fit_csr_counts: integer counts uploaded as uint8 / uint16 and widened on the device give exactly the
    results of the float32 upload (blocking and async, append)."""
    X = synth_csr(180, seed=47)
    top = 255 if dtype == np.uint8 else 60000
    X.data = np.minimum(X.data * (1 if dtype == np.uint8 else 97), top)
    k = 10
    ip, ix = X.indptr.astype(np.int64), np.ascontiguousarray(X.indices.astype(np.uint32))
    f32, narrow = np.ascontiguousarray(X.data.astype(np.float32)), np.ascontiguousarray(X.data.astype(dtype))
    with Fitted(oracle, X, fast=False, n_neighbors=k) as o:
        want = oracle.kneighbors(o.h, k)
    h = cuda.create(n_neighbors=k, fast=False, **REF_KWARGS)
    cuda.fit_csr(h, ip, ix, f32, False)
    ref = cuda.kneighbors(h, k)
    for async_upload in (False, True):
        cuda.fit_csr(h, ip, ix, narrow, False, async_upload=async_upload)
        got = cuda.kneighbors(h, k)
        assert all(np.array_equal(a, b) for a, b in zip(got, ref))
        assert_neighbours_equal_up_to_ties(*got, *want, RTOL)
    cuda.fit_csr(h, ip[:91], ix[:ip[90]], narrow[:ip[90]], False)
    cuda.fit_csr(h, ip[90:] - ip[90], ix[ip[90]:], narrow[ip[90]:], True)
    got = cuda.kneighbors(h, k)
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    cuda.destroy(h)
    # the host class picks the narrow wire format on its own
    from schicexplorer_b200 import MinHash
    assert MinHash._wire_values(X.data).dtype == dtype
    assert MinHash._wire_values(X.data + 0.5).dtype == np.float32


def test_table_prebuilt_during_upload(cuda, oracle):
    """Host fit with the feature-range hint on a batch large enough for the automatic shared-table path: the table is
    built for every id of the range before the ids have landed, the batch then only does lookup + fallback.
    Signatures must equal the oracle's (sampled rows) and the data-dependent table path's (all rows)."""
    cfg = synth.SynthConfig(7, 700)                      # 700 full-length cells: ~1e7 non-zeros, id range 7.5e6 > nnz/2 ...
    ip, ix, dt, _ = synth.generate_csr(cfg, 0, 700, workers=8)
    reps = 2                                             # ... so repeat the cells: 2e7 non-zeros, automatic table path
    ipr = np.concatenate([[0]] + [ip[1:] + r * ip[-1] for r in range(reps)]).astype(np.int64)
    ixr = np.ascontiguousarray(np.tile(ix, reps))
    assert ipr[-1] >= 2 * cfg.chroms.nbins ** 2
    kw = dict(REF_KWARGS)
    h = cuda.create(n_neighbors=5, fast=True, **kw)
    cuda.fit_csr(h, ipr, ixr, None, False)               # no hint: data-dependent table
    sig_plain = cuda.signatures(h, 800)
    cuda.set_feature_range(h, cfg.chroms.nbins ** 2)
    cuda.fit_csr(h, ipr, ixr, None, False)               # hint: prebuilt full-range table
    sig_pre = cuda.signatures(h, 800)
    cuda.destroy(h)
    assert np.array_equal(sig_plain, sig_pre) and np.array_equal(sig_pre[:700], sig_pre[700:])
    rows = np.arange(0, 700, 29)
    sub_ip = np.concatenate([[0], np.cumsum(ip[rows + 1] - ip[rows])]).astype(np.int64)
    sub_ix = np.concatenate([ix[ip[r]:ip[r + 1]] for r in rows])
    ho = oracle.create(n_neighbors=5, fast=True, **kw)
    oracle.fit_csr(ho, sub_ip, sub_ix, None, False)
    assert np.array_equal(oracle.signatures(ho, 800), sig_pre[rows])
    oracle.destroy(ho)


@pytest.mark.parametrize("k,ef", [(3200, 1), (800, 3)])
def test_rerank_long_candidate_lists_are_chunked(cuda, oracle, k, ef):
    """Exact rerank with k * excess_factor beyond what one CTA holds in shared memory (the CLI default k = n on a
    few thousand cells with -em): K4 runs over chunks of 1024 candidates and a per-row merge picks the final k.
    Three clusters of different sizes, so the lists end inside the first, the second and the third chunk."""
    rng = np.random.default_rng(31)
    ncols = 1 << 21
    rows, cols = [], []
    r = 0
    for size in (2400, 700, 100):
        pool = rng.choice(ncols, size=3000, replace=False)
        for _ in range(size):
            c = np.unique(rng.choice(pool, size=240, replace=False))
            rows.append(np.full(len(c), r)); cols.append(c); r += 1
    cc = np.concatenate(cols)
    X = csr_matrix((rng.integers(1, 30, size=len(cc)).astype(np.float64), (np.concatenate(rows), cc)), shape=(r, ncols))
    kw = dict(fast=False, n_neighbors=k, excess_factor=ef, number_of_hash_functions=128, minimal_blocks_in_common=2)
    with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
        a, b = cuda.kneighbors(g.h, k), oracle.kneighbors(o.h, k)
        if ef == 1:
            assert b[2].max() > 2048 and b[2].min() < 1024    # the lists really end in different chunks
        assert_neighbours_equal_up_to_ties(*a, *b, RTOL)
        if ef == 1:
            gi, gc, gd = cuda.kneighbors_graph(g.h, k)
            oi, oc, od = oracle.kneighbors_graph(o.h, k)
            assert np.array_equal(gi, oi)
            np.testing.assert_allclose(np.sort(gd), np.sort(od), rtol=RTOL)
