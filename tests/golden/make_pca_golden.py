"""Regenerates tests/golden/pca_golden.npz: the reference's PCA post-step (scHicClusterMinHash.py:358-366),

    PCA(n_components=min(shape) - 1).fit_transform(graph.todense())

computed with scikit-learn (version recorded in the file) on the kNN graph of the reference's 20-cell fixture
(k = 20, fast distances, the tool's default kwargs; graph from the CPU oracle, stored as float32 like the C ABI
returns it). Pins oracle/pca_np.py and the device PCA independently of the scikit-learn installed at test time."""
import os
import sys

import numpy as np
import sklearn
from scipy.sparse import csr_matrix
from sklearn.decomposition import PCA

ROOT = os.path.join(os.path.dirname(__file__), "..", "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import REF_KWARGS  # noqa: E402
from oracle import build_oracle, oracle_api  # noqa: E402
from schicexplorer_b200.loader import cells_to_csr, load_cells_npz  # noqa: E402

if __name__ == "__main__":
    here = os.path.dirname(__file__)
    cells, chroms = load_cells_npz(os.path.join(here, "fixture20.npz"))
    X, _ = cells_to_csr(cells, chroms)
    build_oracle()
    o = oracle_api()
    h = o.create(n_neighbors=20, fast=True, **REF_KWARGS)
    o.fit_csr(h, X.indptr, X.indices, None, False)
    ip, ix, dt = o.kneighbors_graph(h, 20)
    o.destroy(h)
    dense = np.asarray(csr_matrix((dt, ix, ip), shape=(20, 20)).todense(), dtype=np.float64)
    pca = PCA(n_components=19)
    emb = pca.fit_transform(dense)
    np.savez_compressed(os.path.join(here, "pca_golden.npz"), sklearn_version=np.array(sklearn.__version__),
                        graph_indptr=ip, graph_indices=ix, graph_data=dt, embedding=emb,
                        singular_values=pca.singular_values_, components=pca.components_, mean=pca.mean_)
    print(emb.shape, pca.singular_values_[:4])
