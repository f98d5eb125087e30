"""CPU suite: the scHicCluster command line with the CPU oracle injected behind NearestNeighbors.
Re-expresses the cases of the reference's schicexplorer/test/test_scHicCluster.py:62-250 on the same
test_matrix.scool: output row order = the scool's cell order and exactly 3 distinct labels (what
are_files_equal_clustering, :33-59, checks), for -drm none | knn | pca x kmeans | spectral, --chromosomes, -ic,
-pca and the scatter plot; plus the flag contract of scHicCluster.py:29-124."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from schicexplorer_b200 import NearestNeighbors, scHicCluster, scool

MATRIX = os.path.join(GOLDEN, "test_matrix.scool")

CASES = [
    ("kmeans", ""), ("kmeans", "--chromosomes chr1 chr2"), ("spectral", ""), ("spectral", "-drm knn"),
    ("spectral", "-drm pca"), ("kmeans", "-drm knn"), ("kmeans", "-drm pca"), ("kmeans", "-drm knn -ic"),
    ("kmeans", "-drm knn -pca -dim_pca 10"), ("spectral", "-drm knn -k 7"),
]


@pytest.fixture
def oracle_backed(oracle, monkeypatch):
    monkeypatch.setattr(NearestNeighbors, "_api_factory", staticmethod(lambda: oracle))


def run_cluster(tmp_path, method, extra, plot=False):
    out = str(tmp_path / "clusters.txt")
    args = f"--matrix {MATRIX} --numberOfClusters 3 --clusterMethod {method} --outFileName {out} -t 2 {extra}".split()
    if plot:
        args += f"-csp {tmp_path / 'pca_plot'} --colorMap tab10 --dpi 100 --fontsize 5 --figuresize 15 8".split()
    labels = scHicCluster.main(args)
    rows = [l.split(" ") for l in open(out).read().strip().split("\n")]
    names = scool.list_scool_cells(MATRIX)
    # -drm pca clusters the n-1 leading eigenvectors (scHicCluster.py:190-192), so the reference's own expected files
    # (test-data/scHicCluster/cluster_*_pca.txt) hold n-1 = 19 rows; everything else one row per cell
    assert len(rows) == (len(names) - 1 if "-drm pca" in extra else len(names))
    assert [r[0] for r in rows] == names[:len(rows)]
    assert len({r[1] for r in rows}) == 3
    return labels


@pytest.mark.parametrize("method,extra", CASES)
def test_reference_cases(oracle_backed, tmp_path, method, extra):
    run_cluster(tmp_path, method, extra)


def test_scatter_plot(oracle_backed, tmp_path):
    pytest.importorskip("matplotlib")
    run_cluster(tmp_path, "kmeans", "-drm knn -ic", plot=True)
    assert os.path.getsize(tmp_path / "pca_plot.png") > 0


def test_knn_step_equals_the_reference_call(oracle_backed):
    """The matrix the tool clusters on after -drm knn == sklearn's graph for the reference's call (scHicCluster.py:179-180)."""
    import warnings
    from sklearn.neighbors import NearestNeighbors as SkNN
    from schicexplorer_b200 import loader
    cells, chroms = scool.read_scool(MATRIX)
    X, _ = loader.cells_to_csr(cells, chroms)
    g = scHicCluster.knn_graph(X, 19)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = SkNN(n_neighbors=19, algorithm="ball_tree").fit(X).kneighbors_graph(mode="distance")
    assert g.shape == s.shape == (20, 20)
    np.testing.assert_allclose(np.asarray(g.todense()), np.asarray(s.todense()), rtol=1e-5)


def test_flag_contract():
    p = scHicCluster.parse_arguments()
    a = p.parse_args("-m x.scool -o out.txt".split())
    assert a.numberOfClusters == 12 and a.clusterMethod == "spectral" and a.dimensionReductionMethod == "none"
    assert a.numberOfNearestNeighbors == 100 and a.dimensionsPCA == 20 and a.threads == 8 and a.dpi == 300
    assert a.additionalPCA is False and a.createScatterPlot is None and tuple(a.figuresize) == (15, 6)
    assert a.colorMap == "glasbey_dark" and a.fontsize == 15
    b = p.parse_args("-m x -o y -c 5 -cm kmeans -drm knn -pca -dim_pca 7 -k 30 -ic --chromosomes chr1 chr2 -cct a.txt "
                     "-ccb b.txt -lt t.tex -csp plot -d 50 -t 3".split())
    assert b.numberOfClusters == 5 and b.dimensionReductionMethod == "knn" and b.additionalPCA and b.dimensionsPCA == 7
    assert b.numberOfNearestNeighbors == 30 and b.intraChromosomalContactsOnly and b.chromosomes == ["chr1", "chr2"]
    with pytest.raises(SystemExit):
        p.parse_args("-m x -o y -drm umap".split())
    for flag in ("--help", "--version"):
        with pytest.raises(SystemExit) as e:
            p.parse_args([flag])
        assert e.value.code == 0
