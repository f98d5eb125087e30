"""GPU suite (-m gpu): seeded random small cases, every kernel implementation against the oracle.

Rows are drawn to hit the awkward spots: ids right at the borders of the rerank's 2^19-id windows and of the 16-id
table entries, ids 0 and 2^32-1, empty / single-element rows, duplicated rows (ties everywhere), H not a multiple of
32 or 2, k larger than the number of rows, thresholds that exclude everything. Bit-exact signatures, counts and
neighbour indices; 1e-5 relative distances."""

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from conftest import Fitted
from test_gpu_parity import RTOL, assert_neighbours_equal_up_to_ties

pytestmark = pytest.mark.gpu

K1 = ["auto", "table", "tile", "direct", "chunk"]
K3 = ["packed", "plain"]
K4 = ["window", "hash"]


def random_case(rng):
    n = int(rng.integers(1, 48))
    wide = rng.random() < 0.3                       # ids over the full 32-bit range (hash-table K4, filtered K1)
    ncols = 2 ** 32 if wide else int(rng.choice([1 << 20, 2744 * 2744, (1 << 21) + 5]))
    pool_size = int(rng.integers(8, 4000))
    anchors = np.array([0, 15, 16, 17, (1 << 19) - 1, 1 << 19, (1 << 19) + 1, (1 << 20) - 1, ncols - 1, ncols - 2])
    anchors = anchors[anchors < ncols]
    pool = np.unique(np.concatenate([rng.integers(0, ncols, size=pool_size, dtype=np.uint64).astype(np.int64), anchors]))
    rows, cols, vals = [], [], []
    for r in range(n):
        kind = rng.random()
        if kind < 0.08:
            continue                                # empty row
        if kind < 0.16 and r > 0 and rows:          # duplicate of the previous non-empty row
            m = rows[-1] == rows[-1][0]
            rows.append(np.full(int(m.sum()), r)); cols.append(cols[-1].copy()); vals.append(vals[-1].copy())
            continue
        m = 1 if kind < 0.24 else int(rng.integers(1, min(len(pool), 6000) + 1))
        c = np.sort(rng.choice(pool, size=m, replace=False))
        rows.append(np.full(m, r)); cols.append(c)
        vals.append(rng.integers(1, 300, size=m).astype(np.float64) if rng.random() < 0.7 else rng.random(m) * 9 + 0.1)
    if not rows:
        rows, cols, vals = [np.zeros(1, dtype=int)], [np.array([pool[0]])], [np.ones(1)]
    X = csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, ncols))
    X.sort_indices()
    if wide:
        X.indices = X.indices.astype(np.int64)
    H = int(rng.choice([1, 2, 31, 33, 64, 96, 127, 255, 300]))
    kw = dict(number_of_hash_functions=H, minimal_blocks_in_common=int(rng.choice([1, 1, 2, H // 2 + 1, H + 1])),
              excess_factor=int(rng.choice([1, 1, 2, 5])), max_bin_size=int(rng.choice([100000, 100000, 3, n + 1])),
              fast=bool(rng.random() < 0.4), n_neighbors=int(rng.integers(1, n + 6)),
              hash_width=int(rng.choice([32, 32, 64])), absolute_numbers=bool(rng.random() < 0.2))
    return X, kw


@pytest.mark.parametrize("seed", range(120))
def test_random_case_all_implementations(cuda, oracle, monkeypatch, seed):
    rng = np.random.default_rng(1000 + seed)
    X, kw = random_case(rng)
    k = kw["n_neighbors"]
    with Fitted(oracle, X, **kw) as o:
        want_sig, want_cnt = o.sig(), oracle.collision_counts(o.h)
        want = oracle.kneighbors(o.h, k)
        want_graph = oracle.kneighbors_graph(o.h, k)
    modes = [(K1[seed % 5], K3[seed % 2], K4[(seed // 2) % 2]), ("auto", "packed", "window")]
    for k1, k3, k4 in modes:
        monkeypatch.setenv("SCHIC_K1_MODE", k1)
        monkeypatch.setenv("SCHIC_K3_MODE", k3)
        monkeypatch.setenv("SCHIC_K4_MODE", k4)
        with Fitted(cuda, X, **kw) as g:
            data = None if kw["fast"] else X.data.astype(np.float32)
            if seed % 3 == 0:
                cuda.set_feature_range(g.h, X.shape[1])
                cuda.fit_csr(g.h, X.indptr, X.indices, data, False)
            if seed % 4 == 1 and X.shape[0] > 1:            # the same rows through fit + partial_fit
                cut = X.shape[0] // 2
                a, b = X[:cut], X[cut:]
                cuda.fit_csr(g.h, a.indptr, a.indices, None if kw["fast"] else a.data.astype(np.float32), False)
                cuda.fit_csr(g.h, b.indptr, b.indices, None if kw["fast"] else b.data.astype(np.float32), True)
            assert np.array_equal(g.sig(), want_sig), (k1, k3, k4)
            assert np.array_equal(cuda.collision_counts(g.h), want_cnt), (k1, k3, k4)
            got = cuda.kneighbors(g.h, k)
            assert_neighbours_equal_up_to_ties(*got, *want, 0 if kw["fast"] else RTOL)
            gi, gc, gd = cuda.kneighbors_graph(g.h, k)
            assert np.array_equal(gi, want_graph[0])
            if kw["fast"]:
                assert np.array_equal(gc, want_graph[1]) and np.array_equal(gd, want_graph[2])


@pytest.mark.parametrize("H", [4094, 4095, 4096, 5003])
def test_hash_counts_around_the_packed_and_histogram_limits(cuda, oracle, H):
    """H = 4094 is the last hash count of the packed K3 (fp16 counters), 4095 the last of the single-histogram select;
    above them the 32-bit compare kernel and the two-level radix select take over."""
    rng = np.random.default_rng(H)
    X, kw = random_case(rng)
    kw.update(number_of_hash_functions=H, minimal_blocks_in_common=1, hash_width=32, fast=True, max_bin_size=100000)
    k = kw["n_neighbors"]
    with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
        assert np.array_equal(g.sig(), o.sig())
        assert np.array_equal(cuda.collision_counts(g.h), oracle.collision_counts(o.h))
        a, b = cuda.kneighbors(g.h, k), oracle.kneighbors(o.h, k)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("fast", [True, False])
def test_several_query_blocks(cuda, oracle, monkeypatch, fast):
    """The count matrix is produced in query blocks when it exceeds the memory budget (n beyond ~100 k cells); here the
    budget is forced down so that 300 cells need three blocks: non-symmetric K3 per block, select / rerank offsets."""
    from schicexplorer_b200 import synth
    cfg = synth.small_config(300, config_id=51)
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=1)
    X = synth.to_scipy(cfg, ip, ix, dt)
    monkeypatch.setenv("SCHIC_COUNT_BUDGET", str(128 * 304))
    kw = dict(fast=fast, n_neighbors=17, number_of_hash_functions=200, minimal_blocks_in_common=10)
    with Fitted(cuda, X, **kw) as g, Fitted(oracle, X, **kw) as o:
        a, b = cuda.kneighbors(g.h, 17), oracle.kneighbors(o.h, 17)
        assert_neighbours_equal_up_to_ties(*a, *b, 0 if fast else RTOL)
        assert np.array_equal(cuda.collision_counts(g.h), oracle.collision_counts(o.h))


@pytest.mark.parametrize("width", [32, 64])
def test_ids_beyond_32_bits(cuda, oracle, width):
    """Feature ids >= 2^32 (10 kb resolution: bins^2 ~ 7e10): the width-32 hash uses the low word, the width-64 hash
    both words. Includes partial_fit where only the second batch has large ids (high words are back-filled)."""
    rng = np.random.default_rng(width)
    ncols = 1 << 40
    small = np.unique(rng.integers(0, 1 << 31, size=400))
    large = np.unique(rng.integers(1 << 32, ncols, size=400))
    rows, cols = [], []
    for r in range(30):
        pool = small if r < 12 else np.concatenate([small, large])
        c = np.sort(rng.choice(pool, size=int(rng.integers(1, 300)), replace=False))
        rows.append(np.full(len(c), r)); cols.append(c)
    X = csr_matrix((np.ones(sum(len(c) for c in cols)), (np.concatenate(rows), np.concatenate(cols))), shape=(30, ncols))
    X.indices = X.indices.astype(np.int64)
    X.indptr = X.indptr.astype(np.int64)
    kw = dict(number_of_hash_functions=96, minimal_blocks_in_common=1, hash_width=width, fast=True, n_neighbors=9)
    with Fitted(oracle, X, **kw) as o:
        want_sig, want = o.sig(), oracle.kneighbors(o.h, 9)
    with Fitted(cuda, X, **kw) as g:
        assert np.array_equal(g.sig(), want_sig)
        got = cuda.kneighbors(g.h, 9)
        assert all(np.array_equal(a, b) for a, b in zip(got, want))
        a, b = X[:12], X[12:]
        cuda.fit_csr(g.h, a.indptr, a.indices, None, False)          # small ids only
        cuda.fit_csr(g.h, b.indptr, b.indices, None, True)           # large ids appear: back-fill
        assert np.array_equal(g.sig(), want_sig)


@pytest.mark.parametrize("width", [32, 64])
def test_exact_rerank_with_ids_beyond_32_bits(cuda, oracle, width):
    """-em at high resolution (10 kb: bins^2 ~ 7e10 feature ids): the rerank and the exact kNN run on the ranks of the
    64-bit ids (k7_relabel.cu). Ids that differ only in the HIGH word must stay different features; integer counts take
    the counts kernel, fractional values the generic one; partial_fit re-ranks over all fitted rows; the Python class
    takes the same route for a matrix with more than 2^32 columns."""
    rng = np.random.default_rng(40 + width)
    ncols = 1 << 40
    low = np.unique(rng.integers(0, 1 << 20, size=600))
    pool = np.unique(np.concatenate([low, low + (1 << 32), low + (5 << 32), rng.integers(1 << 33, ncols, size=500)]))
    rows, cols, vals = [], [], []
    for r in range(80):
        c = np.sort(rng.choice(pool, size=int(rng.integers(20, 700)), replace=False))
        rows.append(np.full(len(c), r)); cols.append(c); vals.append(rng.integers(1, 30, size=len(c)).astype(np.float64))
    X = csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(80, ncols))
    X.indices = X.indices.astype(np.int64)
    X.indptr = X.indptr.astype(np.int64)
    for scale, kernel in ((1.0, "k4c_rerank_direct_kernel"), (0.5, "k4_rerank_window_kernel")):
        Xs = X.copy(); Xs.data = Xs.data * scale
        kw = dict(number_of_hash_functions=128, minimal_blocks_in_common=1, hash_width=width, fast=False, n_neighbors=12, excess_factor=2)
        with Fitted(oracle, Xs, **kw) as o, Fitted(cuda, Xs, **kw) as g:
            want = oracle.kneighbors(o.h, 12)
            assert_neighbours_equal_up_to_ties(*cuda.kneighbors(g.h, 12), *want, RTOL)
            assert cuda.rerank_kernel_name(g.h) == kernel
            assert_neighbours_equal_up_to_ties(*cuda.kneighbors_exact(g.h, 7, False), *oracle.kneighbors_exact(o.h, 7, False), RTOL)
            a, b = Xs[:30], Xs[30:]
            cuda.fit_csr(g.h, a.indptr, a.indices, a.data.astype(np.float32), False)
            cuda.fit_csr(g.h, b.indptr, b.indices, b.data.astype(np.float32), True)
            assert_neighbours_equal_up_to_ties(*cuda.kneighbors(g.h, 12), *want, RTOL)
    from schicexplorer_b200 import MinHash
    kwc = dict(number_of_hash_functions=128, minimal_blocks_in_common=1, hash_width=width, fast=False, n_neighbors=12, excess_factor=2,
               max_bin_size=100000)
    mh = MinHash(device=0, **kwc)
    graph = mh.fit(X).kneighbors_graph(mode="distance")
    mh.close()
    with Fitted(oracle, X, **dict(kw, fast=False)) as o:
        oi, oc, od = oracle.kneighbors_graph(o.h, 12)
    assert np.array_equal(graph.indptr, oi) and np.array_equal(graph.indices, oc)
    np.testing.assert_allclose(graph.data, od, rtol=RTOL)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("exchange", ["csr", "csr+packed", "packed"])
def test_sharded_query_path_on_one_gpu(cuda, oracle, seed, exchange):
    """The multi-GPU query path emulated on one device: random contiguous shards, one handle per shard; the 'gathered'
    database (signatures, CSR) is handed back to every handle together with the rerank preparation computed shard by
    shard by its owner (schic_mh_rerank_prep_rows -> schic_mh_set_database_prep_device) and a ready event
    (schic_mh_set_database_ready_event). Every shard must answer exactly like the single-handle build.
    exchange: what the ranks hand over for the exact rerank -- the CSR ids + values (generic kernels), those plus the
    packed stream (counts kernel, generic behind it), or the packed stream ALONE (half the exchange volume)."""
    import torch
    from schicexplorer_b200 import synth
    rng = np.random.default_rng(77 + seed)
    n = int(rng.integers(60, 260))
    cfg = synth.small_config(n, config_id=60 + seed)
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=1)
    X = synth.to_scipy(cfg, ip, ix, dt)
    dev = torch.device("cuda:0")
    k, H = int(rng.integers(3, 30)), 160
    kw = dict(n_neighbors=k, number_of_hash_functions=H, fast=False, minimal_blocks_in_common=20, excess_factor=int(rng.choice([1, 2])),
              max_bin_size=100000)
    with Fitted(oracle, X, **kw) as o:
        ref = oracle.kneighbors(o.h, k)
    cuts = sorted(set([0, n] + rng.integers(1, n, size=int(rng.integers(1, 4))).tolist()))
    d_ip = torch.from_numpy(ip).to(dev)
    d_ix = torch.from_numpy(ix.view(np.int32)).to(dev)
    d_dt = torch.from_numpy(dt).to(dev)
    sig_all = torch.empty((n, H), dtype=torch.int32, device=dev)
    handles, nw = [], None
    norms_all = torch.empty(n, dtype=torch.float64, device=dev)
    packed_all = torch.zeros(int(ip[-1]) + 4, dtype=torch.int32, device=dev) if "packed" in exchange else None
    flags_all = torch.zeros(4, dtype=torch.int32, device=dev)
    for a, b in zip(cuts[:-1], cuts[1:]):
        h = cuda.create(device=0, **kw)
        cuda.set_stream(h, torch.cuda.current_stream().cuda_stream)
        cuda.set_feature_range(h, X.shape[1])
        sub_ip = (d_ip[a:b + 1] - d_ip[a]).contiguous()
        lo, hi = int(ip[a]), int(ip[b])
        cuda.fit_csr_device(h, b - a, hi - lo, sub_ip.data_ptr(), d_ix[lo:hi].data_ptr(), d_dt[lo:hi].data_ptr(), False)
        p, cnt = cuda.signatures_device(h)
        torch.cuda.synchronize()
        sig_all[a:b] = torch.from_numpy(cuda.signatures(h, H).view(np.int32)).to(dev)
        if nw is None:
            nw = cuda.rerank_windows(h)
            assert nw > 0
            dir_all = torch.empty((n, nw + 1), dtype=torch.int32, device=dev)
        flags = torch.empty(4, dtype=torch.int32, device=dev)
        cuda.rerank_prep_rows(h, b - a, sub_ip.data_ptr(), d_ix[lo:hi].data_ptr(), d_dt[lo:hi].data_ptr(), nw,
                              norms_all[a:b].data_ptr(), dir_all[a:b].data_ptr(), flags.data_ptr(),
                              None if packed_all is None else packed_all[lo:hi].data_ptr())
        flags_all += flags          # [0] / [2] are 0/1 flags, [1] a count: a sum combines all of them
        handles.append((a, b, h, sub_ip))
    ev = torch.cuda.Event()
    ev.record()
    for a, b, h, _ in handles:
        csr = exchange != "packed"
        cuda.set_database_device(h, n, a, sig_all.data_ptr(), d_ip.data_ptr(), d_ix.data_ptr() if csr else None,
                                 d_dt.data_ptr() if csr else None)
        cuda.set_database_ready_event(h, ev.cuda_event)
        if packed_all is None:
            cuda.set_database_prep_device(h, norms_all.data_ptr(), dir_all.data_ptr(), nw)
        else:
            cuda.set_database_prep_device(h, norms_all.data_ptr(), dir_all.data_ptr(), nw, packed_all.data_ptr(),
                                          flags_all.data_ptr(), int(ip[-1]))
        idx, dist, nv = cuda.kneighbors(h, k, n_rows=b - a)
        assert_neighbours_equal_up_to_ties(idx, dist, nv, ref[0][a:b], ref[1][a:b], ref[2][a:b], RTOL)
        assert cuda.rerank_kernel_name(h) == ("k4_rerank_window_kernel" if packed_all is None else "k4c_rerank_direct_kernel")
        cuda.destroy(h)


@pytest.mark.parametrize("seed", range(8))
def test_python_classes_chunked_fit(cuda, seed):
    """The reference-facing classes: MinHashClustering.fit with a random pSaveMemory share (fit + partial_fit in
    chunks, scHicClusterMinHash.py:350-353 / :406) gives the same kNN graph as one fit; values that are integer counts
    and values that are not (different wire formats) both go through."""
    from sklearn.cluster import KMeans
    from schicexplorer_b200 import MinHash, MinHashClustering, synth
    rng = np.random.default_rng(300 + seed)
    n = int(rng.integers(24, 90))
    cfg = synth.small_config(n, config_id=80 + seed)
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=1)
    X = synth.to_scipy(cfg, ip, ix, dt)
    if seed % 2:
        X.data = X.data * 0.37 + 0.01                       # not integer: float32 wire format
    fast = bool(seed % 3)
    kw = dict(n_neighbors=int(rng.integers(2, n)), number_of_hash_functions=int(rng.choice([64, 200, 800])), fast=fast,
              minimal_blocks_in_common=5, excess_factor=1, max_bin_size=100000, shingle_size=5, number_of_cores=2)
    one = MinHash(**kw)
    one.fit(X)
    g1 = one.kneighbors_graph(mode="distance")
    share = float(rng.choice([0.1, 0.25, 0.34, 0.5, 0.99]))
    mhc = MinHashClustering(minHashObject=MinHash(**kw), clusteringObject=KMeans(n_clusters=2, n_init=2, random_state=0))
    mhc.fit(X=X, pSaveMemory=share, pPca=False, pPcaDimensions=10, pUmap=False, pUmapDict={})
    g2 = mhc._minHashObject.kneighbors_graph(mode="distance")
    assert np.array_equal(g1.indptr, g2.indptr) and np.array_equal(g1.indices, g2.indices)
    np.testing.assert_allclose(g1.data, g2.data, rtol=0 if fast else 1e-6)
    assert len(mhc.predict(mhc._precomputed_graph)) == n
    conn = one.kneighbors_graph(mode="connectivity")
    assert np.array_equal(conn.indices, g1.indices) and (conn.data == 1).all()
    one.close(); mhc._minHashObject.close()


@pytest.mark.parametrize("H,hint", [(160, False), (160, True), (1500, True), (33, False)])
def test_medium_batches_take_the_table_path_on_their_own(cuda, oracle, H, hint):
    """Batches big enough (>= 2 M non-zeros, id range <= nnz / 2) choose the shared-table K1 without being told;
    with the feature-range hint the host fit pre-builds the table for every id while the ids upload. H = 1500 with
    6 000 non-zeros per row pushes tau down to 8 (more fallbacks), H = 33 is an odd padded unit."""
    rng = np.random.default_rng(H + int(hint))
    n, ncols = 420, 1 << 20
    rows, cols = [], []
    pool = rng.choice(ncols, size=60000, replace=False)
    for r in range(n):
        m = int(rng.integers(4000, 8000))
        c = np.sort(rng.choice(pool, size=m, replace=False))
        rows.append(np.full(m, r)); cols.append(c)
    cc = np.concatenate(cols)
    X = csr_matrix((np.ones(len(cc)), (np.concatenate(rows), cc)), shape=(n, ncols))
    assert X.nnz >= (1 << 21) and ncols * 2 <= X.nnz
    kw = dict(number_of_hash_functions=H, minimal_blocks_in_common=1, fast=True, n_neighbors=12)
    with Fitted(oracle, X, **kw) as o:
        want_sig = o.sig()
        want = oracle.kneighbors(o.h, 12)
    h = cuda.create(**{**kw, "max_bin_size": 100000})
    if hint:
        cuda.set_feature_range(h, ncols)
    cuda.fit_csr(h, X.indptr, X.indices, None, False)
    assert np.array_equal(cuda.signatures(h, H), want_sig)
    got = cuda.kneighbors(h, 12)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    assert cuda.launch_count(h) > 0
    cuda.destroy(h)


def test_more_rows_than_the_packed_k3_takes(cuda, oracle):
    """n = 60 000 > 57 344: the 16-bit renaming does not apply, the 32-bit compare kernel (symmetric, one block) and the
    uint64-key select path for large n must give the oracle's neighbours."""
    rng = np.random.default_rng(60000)
    n, ncols, m = 60000, 1 << 16, 24
    protos = [np.sort(rng.choice(ncols, size=200, replace=False)) for _ in range(400)]
    rows = np.repeat(np.arange(n), m)
    cols = np.concatenate([np.sort(rng.choice(protos[r % 400], size=m, replace=False)) for r in range(n)])
    X = csr_matrix((np.ones(n * m), (rows, cols)), shape=(n, ncols))
    kw = dict(number_of_hash_functions=32, minimal_blocks_in_common=6, fast=True, n_neighbors=8, max_bin_size=100000)
    with Fitted(oracle, X, **kw) as o:
        want = oracle.kneighbors(o.h, 8)
    with Fitted(cuda, X, **kw) as g:
        got = cuda.kneighbors(g.h, 8)
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
