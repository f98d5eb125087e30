"""CPU suite: the scHicClusterMinHash command line (same flags as the reference) end to end, with the
CPU oracle injected as the C-ABI back end. Re-expresses the assertions of the reference's
schicexplorer/test/test_scHicClusterMinHash.py:35-216 -- cell order of the output file and exactly
3 distinct labels -- on the same 20 cells."""
import os

import pytest

from conftest import GOLDEN
from schicexplorer_b200 import MinHash, scHicClusterMinHash

MATRIX = os.path.join(GOLDEN, "fixture20.npz")


@pytest.fixture(autouse=True)
def oracle_backed(oracle, monkeypatch):
    monkeypatch.setattr(MinHash, "_api_factory", staticmethod(lambda: oracle))


def check_output(path, names, n_clusters=3):
    rows = [l.split(" ") for l in open(path).read().strip().split("\n")]
    assert [r[0] for r in rows] == names                       # row order = cell_name_list order
    assert len({r[1] for r in rows}) == n_clusters             # "exactly 3 distinct labels"


def run(tmp_path, extra, method="kmeans"):
    out = str(tmp_path / "clusters.txt")
    args = f"--matrix {MATRIX} --numberOfClusters 3 --clusterMethod {method} --outFileName {out} -t 3 -nh 800 " \
           f"--createScatterPlot {tmp_path / 'plot.png'} {extra}".split()
    scHicClusterMinHash.main(args)
    return out


@pytest.mark.parametrize("method,extra", [
    ("kmeans", "-dim_pca 100"), ("agglomerative_ward", "--noPCA"), ("agglomerative_complete", "--noPCA"),
    ("agglomerative_average", "--noPCA"), ("agglomerative_single", "--noPCA"), ("birch", "-dim_pca 50"),
    ("spectral", ""), ("spectral", "--chromosomes chr1 chr2"), ("kmeans", "-em"), ("spectral", "-em"),
    ("kmeans", "--saveMemory"), ("agglomerative_ward", "--saveMemory --noPCA"), ("spectral", "--saveMemory -em"),
])
def test_cli_reference_cases(tmp_path, fixture20, method, extra):
    out = run(tmp_path, extra, method)
    check_output(out, fixture20[1])


def test_flag_contract():
    p = scHicClusterMinHash.parse_arguments()
    a = p.parse_args("-m x.scool -o out.txt".split())
    # defaults of the reference (scHicClusterMinHash.py:47-240)
    assert a.numberOfClusters == 12 and a.clusterMethod == "spectral" and a.numberOfHashFunctions == 4000
    assert a.numberOfNearestNeighbors is None and a.shareOfMatrixToBeTransferred == 0.25 and a.threads == 8
    assert a.dimensionsPCA == 100 and a.dpi == 300 and a.createScatterPlot == "scatterPlot.eps"
    assert a.umap_n_neighbors == 30 and a.umap_metric == "canberra" and a.umap_min_dist == 0.3
    # the -em polarity trap: store_false, so the default True means fast=True
    assert a.euclideanModeMinHash is True
    assert p.parse_args("-m x -o y -em".split()).euclideanModeMinHash is False
    b = p.parse_args("-m x -o y -c 7 -cm birch -ic -sirm raw.npz -nh 1200 -k 50 -s 0.5 -sm --noPCA --noUMAP "
                     "-dim_pca 20 -d 5000000 --chromosomes chr1 chr2 -av 1 -lt t.tex --runInHyperoptMode -t 4".split())
    assert b.numberOfClusters == 7 and b.intraChromosomalContactsOnly and b.saveMemory and b.distance == 5e6
    assert b.chromosomes == ["chr1", "chr2"] and b.numberOfNearestNeighbors == 50
    with pytest.raises(SystemExit) as e:
        p.parse_args(["--help"])
    assert e.value.code == 0
    with pytest.raises(SystemExit) as e:
        p.parse_args(["--version"])
    assert e.value.code == 0


def test_save_intermediate_raw_matrix(tmp_path, fixture20):
    from scipy.sparse import load_npz
    raw = str(tmp_path / "raw.npz")
    run(tmp_path, f"--noPCA -sirm {raw}", "agglomerative_ward")
    X = load_npz(raw)
    assert X.shape == fixture20[0].shape and X.nnz == fixture20[0].nnz


@pytest.mark.parametrize("fname,prefix", [("test_matrix.scool", "/cells/"), ("test_matrix_old.scool", "/")])
def test_cli_on_real_scool(tmp_path, fname, prefix):
    """The reference's tests pass `--matrix test-data/test_matrix.scool`
    (test_scHicClusterMinHash.py:66-74); same file here, read without cooler/h5py."""
    out = str(tmp_path / "clusters.txt")
    args = f"--matrix {os.path.join(GOLDEN, fname)} --numberOfClusters 3 --clusterMethod kmeans --outFileName {out} " \
           f"-t 3 -nh 800 -dim_pca 100 --createScatterPlot {tmp_path / 'plot.png'}".split()
    scHicClusterMinHash.main(args)
    rows = [l.split(" ") for l in open(out).read().strip().split("\n")]
    assert len(rows) == 20 and all(r[0].startswith(prefix) for r in rows)
    assert len({r[1] for r in rows}) == 3


def _type_file(tmp_path, names):
    path = tmp_path / "types.txt"
    kinds = ["G1", "S", "G2"]
    path.write_text("".join(f"{n}\t{kinds[i % 3]}\n" for i, n in enumerate(names)))
    return str(path)


def test_cell_coloring_type_writes_matches_and_score(tmp_path, fixture20, monkeypatch):
    """-cct: the non-plot outputs of the reference (scHicClusterMinHash.py:519-580) -- matches.txt in the working
    directory and the score file `correct_associated` that scHicClusterMinHashHyperopt reads back."""
    from schicexplorer_b200 import overlap
    monkeypatch.chdir(tmp_path)
    names = fixture20[1]
    out = run(tmp_path, f"-dim_pca 100 -cct {_type_file(tmp_path, names)}")
    labels = [int(l.split(" ")[1]) for l in open(out).read().strip().split("\n")]
    types = overlap.read_cell_types(_type_file(tmp_path, names))
    score = float(open(tmp_path / "correct_associated").read())
    assert score == overlap.correct_associated(labels, [types[n] for n in names])
    lines = [l for l in open(tmp_path / "matches.txt").read().split("\n") if l]
    assert len(lines) == 3 * 3 and all(l.startswith("Computed cluster ") for l in lines)
    # the counts of one line, recomputed
    import re
    m = re.match(r"Computed cluster (\d+) \(size: (\d+)\) matching with cell type (\S+) \(size: (\d+)\) (\d+) times\.", lines[0])
    cluster, kind = int(m.group(1)), m.group(3)
    assert int(m.group(2)) == labels.count(cluster)
    assert int(m.group(5)) == sum(1 for n, lab in zip(names, labels) if lab == cluster and types[n] == kind)


def test_latex_table_replaces_matches(tmp_path, fixture20, monkeypatch):
    monkeypatch.chdir(tmp_path)
    table = tmp_path / "table.tex"
    run(tmp_path, f"-dim_pca 100 -cct {_type_file(tmp_path, fixture20[1])} -lt {table}")
    text = table.read_text()
    assert text.startswith("\\begin{table}[!htb]") and text.rstrip().endswith("\\end{table}")
    assert text.count("\\hline Cluster ") == 1 + 3 and "Correct identified $>80\\%$" in text
    assert not (tmp_path / "matches.txt").exists() and (tmp_path / "correct_associated").exists()


def test_coloring_flags_without_effect_warn(tmp_path, fixture20, caplog):
    import logging
    with caplog.at_level(logging.WARNING):
        run(tmp_path, f"-dim_pca 100 -ccb {_type_file(tmp_path, fixture20[1])} -lt {tmp_path / 't.tex'}")
    text = caplog.text
    assert "-ccb" in text and "-lt/--latexTable needs -cct" in text
    assert not (tmp_path / "t.tex").exists()
