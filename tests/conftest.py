import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (checker). Built on first use with g++."""
    from oracle import build_oracle, oracle_api
    build_oracle()
    return oracle_api()


@pytest.fixture(scope="session")
def cuda():
    """The product back end; fails loudly when the library or the GPU is missing."""
    from schicexplorer_b200 import _capi
    api = _capi.cuda_api()
    assert api.device_count() > 0, "no CUDA device visible: -m gpu tests need a B200"
    return api


@pytest.fixture(scope="session")
def fixture20():
    """The reference's 20-cell test matrix (text dump of test_matrix.scool) as CSR."""
    from schicexplorer_b200.loader import cells_to_csr, load_cells_npz
    cells, chroms = load_cells_npz(os.path.join(GOLDEN, "fixture20.npz"))
    X, names = cells_to_csr(cells, chroms)
    return X, names, cells, chroms


@pytest.fixture(scope="session")
def golden20():
    return np.load(os.path.join(GOLDEN, "golden20.npz"))


# kwargs scHicClusterMinHash passes to MinHash in the default branch (scHicClusterMinHash.py:402-404)
REF_KWARGS = dict(number_of_hash_functions=800, minimal_blocks_in_common=100, excess_factor=1, max_bin_size=100000,
                  shingle_size=5, absolute_numbers=False)


class Fitted:
    """Small helper: one handle on one back end, fitted on a CSR."""

    def __init__(self, api, X, fast=True, n_neighbors=20, **kw):
        params = dict(REF_KWARGS)
        params.update(kw)
        self.api, self.H = api, params["number_of_hash_functions"]
        self.h = api.create(n_neighbors=n_neighbors, fast=fast, **params)
        data = None if fast else X.data.astype(np.float32)
        api.fit_csr(self.h, X.indptr, X.indices, data, False)

    def sig(self):
        return self.api.signatures(self.h, self.H)

    def close(self):
        self.api.destroy(self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class OraclePcaApi:
    """Stands in for the CUDA library behind schicexplorer_b200.pca.PCA in the CPU suite: same method, numpy oracle."""
    backend = "oracle"

    def pca_fit_transform(self, X, n_components, device=-1, want_components=False, return_mean=False):
        from scipy.sparse import issparse
        from oracle import pca_np
        dense = np.asarray(X.todense(), dtype=np.float64) if issparse(X) else np.asarray(X, dtype=np.float64)
        emb, sv, comp = pca_np.pca_fit_transform(dense, n_components)
        self.last_pca_mean = np.where(np.isfinite(dense), dense, 0.0).mean(axis=0) if want_components else None
        out = (emb, sv, (comp if want_components else None))
        return out + (self.last_pca_mean,) if return_mean else out


@pytest.fixture(autouse=True)
def _oracle_pca_in_the_cpu_suite(request, monkeypatch):
    """CPU tests exercise the host logic with the oracles injected behind the C-ABI shells; -m gpu tests run the product."""
    if request.node.get_closest_marker("gpu") is None:
        from schicexplorer_b200.pca import PCA
        monkeypatch.setattr(PCA, "_api_factory", staticmethod(lambda: OraclePcaApi()))
