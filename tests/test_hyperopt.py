"""CPU suite: the hyperopt fast path - host logic with the oracle as the back end.

* the property the fast path rests on, checked on the oracle (which re-hashes from scratch and does NOT assume it):
  switching a fitted handle to H' hash functions == a fresh fit with H' == the first H' signature columns;
* the score of scHicClusterMinHash.py:470-580 on hand-computed cases;
* the tool end to end on the 20-cell fixture: every CSR is hashed once, the result file has the reference's layout."""
import ast
import os

import numpy as np
import pytest

from conftest import GOLDEN, REF_KWARGS, Fitted
from schicexplorer_b200 import MinHash
from schicexplorer_b200 import scHicClusterMinHashHyperopt as hyper

MATRIX = os.path.join(GOLDEN, "fixture20.npz")


@pytest.fixture
def oracle_backed(oracle, monkeypatch):
    monkeypatch.setattr(MinHash, "_api_factory", staticmethod(lambda: oracle))


def test_active_hashes_is_a_prefix_and_equals_a_fresh_fit(oracle, fixture20):
    X = fixture20[0]
    with Fitted(oracle, X, fast=True, n_neighbors=20) as full:          # H = 800
        sig800 = full.sig()
        for H in (300, 800, 1, 799):
            oracle.set_active_hashes(full.h, H)
            got = oracle.signatures(full.h, H)
            assert np.array_equal(got, sig800[:, :H])
            with Fitted(oracle, X, fast=True, n_neighbors=20, **dict(REF_KWARGS, number_of_hash_functions=H)) as fresh:
                assert np.array_equal(got, fresh.sig())
                if H >= 100:
                    assert np.array_equal(oracle.collision_counts(full.h), oracle.collision_counts(fresh.h))
                    for a, b in zip(oracle.kneighbors(full.h, 20), oracle.kneighbors(fresh.h, 20)):
                        assert np.array_equal(a, b)
        with pytest.raises(Exception):
            oracle.set_active_hashes(full.h, 801)
        oracle.set_active_hashes(full.h, 200)
        oracle.fit_csr(full.h, X.indptr, X.indices, None, True)           # partial_fit ends the prefix
        assert oracle.signatures(full.h, 800).shape == (40, 800)
        assert np.array_equal(oracle.signatures(full.h, 800)[:20], sig800)


def test_correct_associated_known_answers():
    types = ["a"] * 4 + ["b"] * 4 + ["c"] * 2
    assert hyper.correct_associated([0] * 4 + [1] * 4 + [2] * 2, types) == 1.0
    assert hyper.correct_associated([0] * 10, types) == 0.0                     # one cluster, no type reaches 80 %
    # cluster 0 = 4 a + 1 b (80 % a -> 4 a detected), cluster 1 = 3 b + 2 c (60 % b -> nothing)
    labels = [0, 0, 0, 0, 0, 1, 1, 1, 1, 1]
    assert hyper.correct_associated(labels, types) == pytest.approx((4 / 4 + 0 + 0) / 3)
    # two pure clusters of the same type both count
    assert hyper.correct_associated([0, 0, 1, 1, 2, 2, 2, 2, 3, 3], types) == 1.0


def test_read_cell_types(tmp_path):
    p = tmp_path / "types.txt"
    p.write_text("cell1\tG1\ncell2    S\n\ncell3\tG1\n")
    assert hyper.read_cell_types(str(p)) == {"cell1": "G1", "cell2": "S", "cell3": "G1"}


def test_resident_cells_hash_once(oracle_backed, oracle, fixture20):
    res = hyper.ResidentCells(MATRIX, 900, 1000)
    assert res.n_neighbors == 20
    for intra, H in [(False, 300), (False, 900), (True, 500), (False, 300), (True, 900)]:
        g = res.graph(intra, H)
        from schicexplorer_b200 import loader
        X, _ = loader.cells_to_csr(res.cells, res.chroms, intra_only=intra)
        with Fitted(oracle, X, fast=True, n_neighbors=20, **dict(REF_KWARGS, number_of_hash_functions=H)) as fresh:
            ip, ix, dt = oracle.kneighbors_graph(fresh.h, 20)
        assert np.array_equal(g.indptr, ip) and np.array_equal(g.indices, ix)
        np.testing.assert_array_equal(g.data.astype(np.float32), np.where(np.isfinite(dt), dt, 0))
    assert res.fits == 2
    res.close()


def test_tool_end_to_end(oracle_backed, fixture20, tmp_path):
    names = fixture20[1]
    colors = tmp_path / "cell_types.txt"
    colors.write_text("".join(f"{n}\t{'ABC'[i % 3]}\n" for i, n in enumerate(names)))
    out = tmp_path / "best.txt"
    args = f"-m {MATRIX} -c {colors} -o {out} --runs 5 -k 1000 -noh 200 900 200 -noc 2 5 -nop 5 16 5 -unon 5 8 1 " \
           f"-unoc 2 4 1 -umin 0.0 0.5 -me random -t 2".split()
    best = hyper.main(args)
    text = out.read_text()
    assert text.startswith("# Created by scHiCExplorer schicClusterMinHashHyperopt ")
    assert ast.literal_eval(text.split("\n\n", 1)[1]) == best
    assert best["numberOfHashFunctions"] in (200, 400, 600, 800) and best["numberOfClusters"] in (2, 3, 4)
    assert best["clusterMethod"] == "spectral" and 0.0 <= best["umap_min_dist"] <= 0.5


def test_flag_contract():
    p = hyper.parse_arguments()
    a = p.parse_args("-m x.scool -c types.txt -me tpe".split())
    assert a.outputFileName == "hyperoptMinHash_result.txt" and a.runs == 100 and a.nearestNeighbors == 1000
    assert tuple(a.numberOfHashfunctions) == (1000, 20000, 1000) and tuple(a.numberOfClusters) == (6, 15)
    assert tuple(a.numberPCADimensions) == (30, 60, 1) and tuple(a.umap_numberOfNeighbors) == (30, 60, 1)
    assert tuple(a.umap_n_components) == (2, 10, 1) and tuple(a.umap_min_dist) == (0.0, 0.5) and a.threads == 16
    with pytest.raises(SystemExit):
        p.parse_args("-m x -c y -me grid".split())


def test_trial_umap_arguments_start_from_the_tool_defaults():
    """The reference objective runs scHicClusterMinHash.main, i.e. with every --umap_* default of the tool (notably
    --umap_metric canberra, scHicClusterMinHash.py:165-168) and only the searched values overridden."""
    d = hyper.trial_umap_dict({'umap_n_neighbors': 41, 'umap_n_components': 3, 'umap_min_dist': 0.11})
    assert d['umap_metric'] == 'canberra' and d['umap_init'] == 'spectral' and d['umap_random'] == 42
    assert d['umap_n_neighbors'] == 41 and d['umap_n_components'] == 3 and d['umap_min_dist'] == 0.11
    assert d['umap_negative_sample_rate'] == 5 and d['umap_target_metric'] == 'categorical'


def test_objective_hands_the_tool_defaults_to_the_embedding(oracle_backed, monkeypatch):
    seen = {}

    def fake_post(graph, pca, dims, use_umap, umap_dict):
        seen.update(umap_dict=umap_dict, use_umap=use_umap)
        return np.asarray(graph.todense())

    from schicexplorer_b200.clustering import MinHashClustering
    monkeypatch.setattr(MinHashClustering, "_post_process", staticmethod(fake_post))
    monkeypatch.setattr(hyper, "_umap_available", lambda: True)
    resident = hyper.ResidentCells(MATRIX, 800, 20, threads=2)
    types = ["a" if i % 2 else "b" for i in range(20)]
    params = {'dimensionsPCA': 10, 'umap_n_components': 2, 'umap_n_neighbors': 5, 'umap_min_dist': 0.2, 'noPCA': False,
              'intraChromosomalContactsOnly': False, 'numberOfHashFunctions': 800, 'numberOfClusters': 2, 'threads': 1}
    hyper.objective(params, resident, types)
    resident.close()
    assert seen['use_umap'] and seen['umap_dict']['umap_metric'] == 'canberra'
