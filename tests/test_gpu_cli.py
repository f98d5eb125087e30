"""GPU suite (-m gpu): the scHicClusterMinHash command line end to end on the CUDA back end, on the reference's
own test_matrix.scool -- the assertions of schicexplorer/test/test_scHicClusterMinHash.py:35-216 (cell order of the
output file, exactly 3 distinct labels), plus: the kNN graph the tool clusters on equals the oracle's."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from schicexplorer_b200 import MinHash, scHicClusterMinHash, scool

pytestmark = pytest.mark.gpu
MATRIX = os.path.join(GOLDEN, "test_matrix.scool")


@pytest.mark.parametrize("method,extra", [
    ("kmeans", "-dim_pca 100"), ("spectral", ""), ("agglomerative_ward", "--noPCA"), ("birch", "-dim_pca 50"),
    ("spectral", "--chromosomes chr1 chr2"), ("kmeans", "-em"), ("kmeans", "--saveMemory"), ("spectral", "--saveMemory -em"),
    ("kmeans", "-ic -d 30000000"),
])
def test_cli_on_gpu(cuda, tmp_path, method, extra):
    out = str(tmp_path / "clusters.txt")
    args = f"--matrix {MATRIX} --numberOfClusters 3 --clusterMethod {method} --outFileName {out} -t 3 -nh 800 " \
           f"--createScatterPlot {tmp_path / 'plot.png'} {extra}".split()
    scHicClusterMinHash.main(args)
    rows = [l.split(" ") for l in open(out).read().strip().split("\n")]
    assert [r[0] for r in rows] == scool.list_scool_cells(MATRIX)
    assert len({r[1] for r in rows}) == 3


def test_cli_graph_equals_oracle(cuda, oracle, fixture20):
    """Same kwargs as the CLI's default branch (scHicClusterMinHash.py:402-404): graph from the GPU == graph from the oracle."""
    X = fixture20[0]
    kw = dict(n_neighbors=20, number_of_hash_functions=800, number_of_cores=4, shingle_size=5, fast=True,
              maxFeatures=int(max(X.getnnz(1))), absolute_numbers=False, max_bin_size=100000,
              minimal_blocks_in_common=100, excess_factor=1, prune_inverse_index=False)
    g = MinHash(**kw)
    g.fit(X)
    assert g._api.backend == "cuda"
    gg = g.kneighbors_graph(mode="distance")
    g.close()
    # the oracle through the C ABI directly
    from conftest import Fitted
    with Fitted(oracle, X, fast=True, n_neighbors=20) as f:
        oi, oc, od = oracle.kneighbors_graph(f.h, 20)
    assert np.array_equal(gg.indptr, oi) and np.array_equal(gg.indices, oc) and np.array_equal(gg.data, od)


@pytest.mark.parametrize("method,extra", [("spectral", "-drm knn"), ("kmeans", "-drm knn"), ("kmeans", "-drm knn -ic"),
                                          ("kmeans", "-drm knn -pca -dim_pca 10"), ("spectral", "-drm knn -k 7")])
def test_schiccluster_knn_cli_on_gpu(cuda, tmp_path, method, extra):
    """scHicCluster -drm knn (reference cases test_scHicCluster.py:100-135, 219-235) with the kNN step on the GPU."""
    from schicexplorer_b200 import scHicCluster
    out = str(tmp_path / "clusters.txt")
    args = f"--matrix {MATRIX} --numberOfClusters 3 --clusterMethod {method} --outFileName {out} -t 2 {extra}".split()
    scHicCluster.main(args)
    rows = [l.split(" ") for l in open(out).read().strip().split("\n")]
    assert [r[0] for r in rows] == scool.list_scool_cells(MATRIX)
    assert len({r[1] for r in rows}) == 3


def test_schiccluster_knn_graph_equals_sklearn(cuda):
    import warnings
    from sklearn.neighbors import NearestNeighbors as SkNN
    from schicexplorer_b200 import loader, scHicCluster
    cells, chroms = scool.read_scool(MATRIX)
    X, _ = loader.cells_to_csr(cells, chroms)
    g = scHicCluster.knn_graph(X, 19)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = SkNN(n_neighbors=19, algorithm="ball_tree").fit(X).kneighbors_graph(mode="distance")
    np.testing.assert_allclose(np.asarray(g.todense()), np.asarray(s.todense()), rtol=1e-5)
