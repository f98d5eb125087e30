"""CPU suite: host-side mirror of the reference interface (MinHash / MinHashClustering / loader),
with the CPU oracle injected as the C-ABI back end so the logic runs without a GPU."""
import numpy as np
import pytest
from scipy.sparse import csr_matrix

from conftest import REF_KWARGS
from schicexplorer_b200 import MinHash, MinHashClustering, loader


@pytest.fixture()
def oracle_backed(oracle, monkeypatch):
    monkeypatch.setattr(MinHash, "_api_factory", staticmethod(lambda: oracle))
    return oracle


def test_import_shim_resolves():
    import sparse_neighbors_search as sns
    assert sns.MinHash is MinHash and sns.MinHashClustering is MinHashClustering


def test_minhash_reference_call_sequence(oracle_backed, fixture20, golden20):
    X = fixture20[0]
    # scHicClusterMinHash.py:402-404 with k = ncells
    mh = MinHash(n_neighbors=20, number_of_hash_functions=800, number_of_cores=4, shingle_size=5, fast=True,
                 maxFeatures=int(max(X.getnnz(1))), absolute_numbers=False, max_bin_size=100000,
                 minimal_blocks_in_common=100, excess_factor=1, prune_inverse_index=False)
    mh.fit(X)
    g = mh.kneighbors_graph(mode="distance")
    assert g.shape == (20, 20) and g.dtype == np.float64
    assert np.array_equal(np.diff(g.indptr), golden20["nv_fast"])
    dist, idx = mh.kneighbors()
    assert np.array_equal(idx, golden20["idx_fast"])
    # what the caller does next (:356-357)
    g = np.nan_to_num(g)
    g.data[np.isinf(g.data)] = 0
    assert np.isfinite(g.todense()).all()


def test_partial_fit_numbers_rows_in_append_order(oracle_backed, fixture20, golden20):
    X = fixture20[0]
    # the --saveMemory branch (:347-353): shingle_size=0, fewer kwargs, fit then partial_fit
    mh = MinHash(n_neighbors=20, number_of_hash_functions=800, number_of_cores=2, shingle_size=0, fast=True,
                 maxFeatures=60000, absolute_numbers=False, minimal_blocks_in_common=100, max_bin_size=100000,
                 excess_factor=1)
    step = int(20 * 0.25)
    for j, i in enumerate(range(0, 20, step)):
        chunk = X[i:i + step]
        mh.fit(chunk) if j == 0 else mh.partial_fit(X=chunk)
    assert mh.n_fitted_ == 20
    assert np.array_equal(mh.signatures(), golden20["sig32"])
    assert np.array_equal(mh.kneighbors(return_distance=False), golden20["idx_fast"])


def test_block_combine_option_through_the_class(oracle_backed, fixture20):
    """``shingle=1`` (upstream's separate switch; the reference does not pass it, so the default path is unchanged): the
    class reports one signature column per table, the graph is built on the folded keys, the hyperopt prefix switch and
    the sharded layer refuse the combination."""
    from oracle import oracle_np
    from schicexplorer_b200.dist import ShardedMinHash
    X = fixture20[0]
    kw = dict(n_neighbors=20, number_of_hash_functions=800, shingle_size=5, fast=True, minimal_blocks_in_common=2,
              excess_factor=1, max_bin_size=100000)
    plain = MinHash(**kw).fit(X)
    folded = MinHash(shingle=1, **kw).fit(X)
    assert plain.signatures().shape == (20, 800) and folded.signatures().shape == (20, 160)
    assert np.array_equal(folded.signatures(), oracle_np.combine(plain.signatures(), 5))
    g = folded.kneighbors_graph(mode="distance")
    cnt = oracle_np.collision_counts(folded.signatures())
    assert np.array_equal(np.diff(g.indptr), (cnt >= 2).sum(axis=1))
    with pytest.raises(ValueError):
        folded.set_number_of_hash_functions(400)
    assert folded.get_params()["shingle"] == 1
    import torch
    with pytest.raises(NotImplementedError):
        ShardedMinHash(oracle_backed, 20, [0, 20], torch.device("cpu"), shingle=1, **kw)


def test_unknown_kwargs_warn_but_work(oracle_backed):
    with pytest.warns(UserWarning):
        MinHash(n_neighbors=3, some_new_upstream_flag=1)
    MinHash(n_neighbors=3, gpu_hashing=0, cpu_gpu_load_balancing=0)   # known upstream kwargs: silent


def test_dense_and_unsorted_input(oracle_backed):
    rng = np.random.default_rng(1)
    D = (rng.random((6, 50)) < 0.3) * rng.integers(1, 5, (6, 50))
    a = MinHash(n_neighbors=3, number_of_hash_functions=32, fast=False).fit(D)
    b = MinHash(n_neighbors=3, number_of_hash_functions=32, fast=False).fit(csr_matrix(D))
    assert np.array_equal(a.signatures(), b.signatures())
    da, ia = a.kneighbors()
    db, ib = b.kneighbors()
    assert np.array_equal(ia, ib) and np.array_equal(da, db)


@pytest.mark.parametrize("method", ["kmeans", "agglomerative_ward", "spectral"])
def test_clustering_reference_assertions(oracle_backed, fixture20, method):
    """Re-expression of schicexplorer/test/test_scHicClusterMinHash.py:66-140: the reference asserts
    the cell order of the output and that exactly 3 distinct labels come out."""
    from sklearn.cluster import AgglomerativeClustering, KMeans, SpectralClustering
    X, names = fixture20[0], fixture20[1]
    obj = {"kmeans": KMeans(n_clusters=3, random_state=0, n_init=10),
           "agglomerative_ward": AgglomerativeClustering(n_clusters=3, linkage="ward"),
           "spectral": SpectralClustering(n_clusters=3, affinity="nearest_neighbors", n_neighbors=10, random_state=0)}[method]
    mh = MinHash(n_neighbors=20, fast=True, **REF_KWARGS)
    mhc = MinHashClustering(minHashObject=mh, clusteringObject=obj)
    mhc.fit(X=X, pSaveMemory=0.25, pPca=True, pPcaDimensions=100, pUmap=False, pUmapDict={})
    assert mhc._precomputed_graph.shape == (20, 19)        # PCA(n-1) on the dense 20x20 graph
    labels = mhc.predict(mhc._precomputed_graph, pPca=False, pPcaDimensions=100)
    assert len(labels) == len(names) == 20 and len(set(labels.tolist())) == 3


def test_clustering_without_pca_keeps_sparse_graph(oracle_backed, fixture20):
    from sklearn.cluster import AgglomerativeClustering
    mh = MinHash(n_neighbors=20, fast=True, **REF_KWARGS)
    mhc = MinHashClustering(minHashObject=mh, clusteringObject=AgglomerativeClustering(n_clusters=3, linkage="single"))
    mhc.fit(X=fixture20[0], pSaveMemory=1.0, pPca=False, pUmap=False)
    assert mhc._precomputed_graph.shape == (20, 20)
    # the caller's scrub (:409-413) works on .data
    mhc._precomputed_graph.data[np.isnan(mhc._precomputed_graph.data)] = 0
    assert len(set(mhc.predict(mhc._precomputed_graph).tolist())) == 3


# ---- loader semantics (utilities.py) -----------------------------------------------------------
def test_loader_feature_encoding_and_order(fixture20):
    X, names, cells, chroms = fixture20
    assert chroms.nbins == 2744 and X.shape == (20, 2744 * 2744) and X.nnz == 342382
    assert names[0].startswith("/cells/Diploid_1_CGTACTAG_AAGGAGTA")
    c0 = cells[0]
    assert X[0].nnz == len(c0.bin1) == 55498                  # docs/content/Example_analysis.rst:296
    f = np.sort(c0.bin1 * 2744 + c0.bin2)                     # utilities.py:119-120
    assert np.array_equal(X[0].indices, f)
    assert X.has_sorted_indices and X.dtype == np.float64


def test_loader_filters(fixture20):
    _, _, cells, chroms = fixture20
    c = cells[0]
    # --chromosomes chr1 chr2: pixels whose bin1 lies in chr1 or chr2 (full-genome feature ids kept)
    b1, b2, cnt = loader.select_pixels(c, chroms, chromosomes=["chr1", "chr2"])
    ch = chroms.chrom_of_bin(b1)
    assert set(np.unique(ch)) <= {chroms.names.index("chr1"), chroms.names.index("chr2")}
    assert len(b1) == 8147                                     # SURVEY.md 8(d) config 5 figure
    # -ic keeps bin2 in the same chromosome
    b1, b2, _ = loader.select_pixels(c, chroms, intra_only=True)
    assert (chroms.chrom_of_bin(b1) == chroms.chrom_of_bin(b2)).all() and 0 < len(b1) < len(c.bin1)
    # -d: |bin1-bin2| < d and the scalar d > 5 test (utilities.py:106-113)
    b1, b2, _ = loader.select_pixels(c, chroms, distance_bins=10)
    assert (np.abs(b1 - b2) < 10).all() and len(b1) > 0
    assert len(loader.select_pixels(c, chroms, distance_bins=5)[0]) == 0
    X, _ = loader.cells_to_csr(cells[:3], chroms, chromosomes=["chr1", "chr2"])
    assert X.shape == (3, 2744 * 2744)


def test_empty_cell_keeps_its_row(fixture20):
    _, _, cells, chroms = fixture20
    empty = loader.CellPixels("/cells/empty", np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0))
    X, names = loader.cells_to_csr([cells[1], empty, cells[2]], chroms)
    assert X.shape[0] == 3 and X[1].nnz == 0 and names[1] == "/cells/empty"


def test_wire_values_pick_the_narrowest_exact_format():
    """MinHash._wire_values: integer contact counts cross the C ABI as uint8 / uint16, everything else as float32."""
    from schicexplorer_b200 import MinHash
    w = MinHash._wire_values
    assert w(np.array([1.0, 2.0, 255.0])).dtype == np.uint8
    assert w(np.array([1.0, 2.0, 256.0])).dtype == np.uint16
    assert w(np.array([0, 7, 65535], dtype=np.int64)).dtype == np.uint16
    assert w(np.array([1.0, 65536.0])).dtype == np.float32
    assert w(np.array([1.0, 2.5])).dtype == np.float32
    assert w(np.array([-1.0, 2.0])).dtype == np.float32
    assert w(np.array([1.0, np.nan])).dtype == np.float32
    assert w(np.array([], dtype=np.float64)).dtype == np.float32
    assert np.array_equal(w(np.array([3.0, 4.0])), [3, 4])


def test_loader_matches_the_coo_constructor_on_unsorted_and_duplicated_pixels(fixture20):
    """cells_to_csr builds the CSR row by row; the result must be what utilities.py:131's COO -> CSR constructor gives
    (duplicates summed, columns sorted, scipy's index dtype), also for pixel tables that are not sorted / not unique
    and for --chromosomes given in another order than the file's."""
    from scipy.sparse import csr_matrix
    from schicexplorer_b200 import loader
    from schicexplorer_b200.loader import CellPixels
    _, _, cells, chroms = fixture20
    rng = np.random.default_rng(4)
    messy = []
    for i, c in enumerate(cells[:8]):
        p = rng.permutation(len(c.bin1))
        dup = rng.integers(0, len(c.bin1), size=7)
        b1 = np.concatenate([c.bin1[p], c.bin1[dup]]); b2 = np.concatenate([c.bin2[p], c.bin2[dup]])
        cnt = np.concatenate([c.count[p], c.count[dup] + i])
        messy.append(CellPixels(c.name, b1, b2, cnt))
    messy.append(CellPixels("/cells/empty", np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0)))
    for kw in (dict(), dict(intra_only=True), dict(chromosomes=["chr2", "chr1"]), dict(distance=30_000_000)):
        X, names = loader.cells_to_csr(messy, chroms, **kw)
        rows, feats, vals = [], [], []
        dbins = None if "distance" not in kw else kw["distance"] // chroms.binsize
        for r, c in enumerate(messy):
            b1, b2, cnt = loader.select_pixels(c, chroms, kw.get("chromosomes"), kw.get("intra_only", False), dbins)
            rows.append(np.full(len(b1), r)); feats.append(b1 * np.int64(chroms.nbins) + b2); vals.append(cnt)
        want = csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(feats))),
                          shape=(len(messy), chroms.nbins ** 2), dtype=np.float64)
        want.sort_indices()
        assert X.shape == want.shape and X.indices.dtype == want.indices.dtype and X.has_sorted_indices
        assert np.array_equal(X.indptr, want.indptr) and np.array_equal(X.indices, want.indices)
        assert np.array_equal(X.data, want.data) and names[-1] == "/cells/empty"


def test_tools_and_entry_points_parse():
    """Every script under tools/ and bin/, bench.py and __graft_entry__.py at least parses (they only run on the GPU box)."""
    import ast
    import glob
    import os
    from conftest import ROOT
    files = glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "bin", "*"))
    files += [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    assert len(files) > 10
    for f in files:
        ast.parse(open(f).read(), filename=f)


def test_reconcile_upstream_tool(tmp_path, oracle, monkeypatch, capsys):
    """tools/reconcile_upstream.py: (1) today -- no real sparse_neighbors_search anywhere -- it says so and exits 3, and
    it does not mistake this repository's import shim for upstream; (2) the day an install appears under a path it is
    given, it runs the reference's call sequence on the 20-cell fixture and pins the HashSpec: exercised here with a
    stand-in package whose MinHash is the oracle-backed class (so width 32 must reproduce the golden vectors, width 64
    must not)."""
    import importlib.util
    import json
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("reconcile_upstream", os.path.join(ROOT, "tools", "reconcile_upstream.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    assert tool.main(["--json"]) == 3
    assert json.loads(capsys.readouterr().out)["pinned"] is False
    mod, why = tool.find_upstream(os.path.join(ROOT))       # the shim's parent directory: must be refused
    assert mod is None

    pkg = tmp_path / "sparse_neighbors_search"
    pkg.mkdir()
    (pkg / "__init__.py").write_text(
        "from schicexplorer_b200 import MinHash as _M\n"
        "from oracle import oracle_api\n"
        "__version__ = 'stand-in'\n"
        "class MinHash(_M):\n"
        "    _api_factory = staticmethod(oracle_api)\n")
    assert tool.main(["--json", "--path", str(tmp_path)]) == 0
    rep = json.loads(capsys.readouterr().out)
    assert rep["pinned"] and rep["hash_width_matching_upstream"] == [32]
    assert rep["checks"]["fast_counts_width32"]["equal"] and not rep["checks"]["fast_counts_width64"]["equal"]
    assert rep["checks"]["signatures_width32"]["equal"] and rep["checks"]["exact_distances"]["identical_up_to_ties"]
