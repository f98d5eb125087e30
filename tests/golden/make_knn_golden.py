"""Regenerates tests/golden/knn_golden.npz: the exact-kNN graph the reference's scHicCluster -drm knn step produces
(scHicCluster.py:179-180), computed by the very call the reference makes,

    NearestNeighbors(n_neighbors=k, algorithm='ball_tree').fit(X).kneighbors_graph(mode='distance')

with scikit-learn (the reference's dependency, importable in the build container; version recorded in the file) on
(a) the reference's 20-cell fixture and (b) a seeded synthetic 150-cell matrix. Unlike the MinHash path this row of
SURVEY 8(f) is therefore PINNED to the reference's own arithmetic: tests/test_exact_knn.py holds the oracle to these
vectors on CPU, the GPU suite holds the CUDA path to the oracle and to these vectors."""
import os
import sys
import warnings

import numpy as np
import sklearn
from sklearn.neighbors import NearestNeighbors

ROOT = os.path.join(os.path.dirname(__file__), "..", "..")
sys.path.insert(0, ROOT)
from schicexplorer_b200 import synth  # noqa: E402
from schicexplorer_b200.loader import cells_to_csr, load_cells_npz  # noqa: E402


def synthetic150():
    cfg = synth.small_config(150, config_id=41)
    ip, ix, dt, _ = synth.generate_csr(cfg, workers=1)
    return synth.to_scipy(cfg, ip, ix, dt)


def reference_graph(X, k):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")      # "cannot use tree with sparse input: using brute force"
        nbrs = NearestNeighbors(n_neighbors=k, algorithm="ball_tree").fit(X)
        g = nbrs.kneighbors_graph(mode="distance").tocsr()
    g.sort_indices()
    return g


if __name__ == "__main__":
    here = os.path.dirname(__file__)
    cells, chroms = load_cells_npz(os.path.join(here, "fixture20.npz"))
    X20, _ = cells_to_csr(cells, chroms)
    out = {"sklearn_version": np.array(sklearn.__version__)}
    for tag, X, ks in (("f20", X20, (5, 19)), ("s150", synthetic150(), (10,))):
        for k in ks:
            g = reference_graph(X, k)
            out[f"{tag}_k{k}_indptr"] = g.indptr.astype(np.int64)
            out[f"{tag}_k{k}_indices"] = g.indices.astype(np.int32)
            out[f"{tag}_k{k}_data"] = g.data.astype(np.float64)
    np.savez_compressed(os.path.join(here, "knn_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})
