"""Drop-in for ``sparse_neighbors_search.MinHash`` as scHiCExplorer uses it.

Boundary mirrored: constructor kwargs at
``/root/reference/schicexplorer/scHicClusterMinHash.py:347-348`` and ``:402-404``; methods
``fit`` (:351), ``partial_fit`` (:353), ``kneighbors_graph(mode='distance')`` (:355). The class is
a thin host shell: every number it returns is produced by the sm_100a kernels behind the C ABI of
``include/schic_mh.h``. There is no CPU fallback -- constructing the back end raises if the CUDA
library is not built or no B200 is visible.
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy.sparse import csr_matrix, issparse

from . import _capi

# kwargs of upstream's constructor that have no effect on this path (accepted, ignored)
_IGNORED_KWARGS = {
    "radius", "similarity", "chunk_size", "prune_inverse_index_after_instance",
    "remove_hash_function_with_less_entries_as", "block_size", "store_value_with_least_sigificant_bit",
    "cpu_gpu_load_balancing", "gpu_hashing", "speed_optimized", "accuracy_optimized", "maxFeatures",
    "rangeK_wta",
}


class MinHash:
    """Approximate k-nearest-neighbour search on sparse rows with MinHash signatures.

    Parameters follow upstream; the ones scHiCExplorer passes are ``n_neighbors``,
    ``number_of_hash_functions``, ``number_of_cores``, ``shingle_size``, ``fast``, ``maxFeatures``,
    ``absolute_numbers``, ``max_bin_size``, ``minimal_blocks_in_common``, ``excess_factor``,
    ``prune_inverse_index``. ``hash_width`` (32 default, or 64) selects the HashSpec of
    SURVEY.md 8(c) rule 2; ``device`` the CUDA ordinal.
    """

    # the C-ABI back end; tests swap this for the CPU oracle to exercise the host logic without a GPU
    _api_factory = staticmethod(_capi.cuda_api)

    def __init__(self, n_neighbors=5, fast=False, number_of_hash_functions=400, max_bin_size=50,
                 minimal_blocks_in_common=1, shingle_size=4, excess_factor=5, number_of_cores=None,
                 prune_inverse_index=-1, shingle=0, absolute_numbers=False, hash_width=32, device=-1, **kwargs):
        unknown = set(kwargs) - _IGNORED_KWARGS
        if unknown:
            warnings.warn(f"MinHash: ignoring unsupported keyword arguments {sorted(unknown)}")
        if prune_inverse_index not in (False, None, -1, 0):
            warnings.warn("MinHash: prune_inverse_index is not supported on this path; ignoring")
        self.n_neighbors = int(n_neighbors)
        self.fast = bool(fast)
        self.number_of_hash_functions = int(number_of_hash_functions)
        self.max_bin_size = int(max_bin_size)
        self.minimal_blocks_in_common = int(minimal_blocks_in_common)
        self.shingle_size = int(shingle_size)
        self.excess_factor = int(excess_factor)
        self.number_of_cores = number_of_cores
        self.shingle = int(shingle)
        self.absolute_numbers = bool(absolute_numbers)
        self.hash_width = int(hash_width)
        self.device = int(device)
        self._api = None
        self._h = None
        self._n_features = None

    # ---- life cycle ------------------------------------------------------------------------
    def _ensure(self):
        if self._h is None:
            self._api = self._api_factory()
            self._n_hashes_created = self.number_of_hash_functions
            self._h = self._api.create(
                n_neighbors=self.n_neighbors, number_of_hash_functions=self.number_of_hash_functions,
                hash_width=self.hash_width, fast=self.fast, absolute_numbers=self.absolute_numbers,
                minimal_blocks_in_common=self.minimal_blocks_in_common, excess_factor=self.excess_factor,
                max_bin_size=self.max_bin_size, shingle_size=self.shingle_size, shingle=self.shingle,
                number_of_cores=self.number_of_cores or 0, device=self.device)
        return self._api, self._h

    def close(self):
        if self._h is not None:
            self._api.destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- fit ---------------------------------------------------------------------------------
    @staticmethod
    def _as_csr(X):
        if not issparse(X):
            X = csr_matrix(np.asarray(X))
        X = X.tocsr()
        if not X.has_sorted_indices:
            X = X.sorted_indices()
        return X

    @staticmethod
    def _wire_values(d: np.ndarray) -> np.ndarray:
        """Values as they cross the C ABI: Hi-C values are integer contact counts, which go down as uint8 / uint16
        (a quarter / half of the PCIe bytes; widened to float32 on the device, bit-identical results), anything
        else as float32."""
        if d.size and d.dtype.kind in "fiu":
            lo, hi = d.min(), d.max()
            if lo >= 0 and hi <= 65535:
                narrow = d.astype(np.uint8 if hi <= 255 else np.uint16)
                if d.dtype.kind in "iu" or np.array_equal(narrow, d):
                    return narrow
        return d.astype(np.float32, copy=False)

    def _fit(self, X, append: bool):
        X = self._as_csr(X)
        if append and self._n_features is not None and X.shape[1] != self._n_features:
            raise ValueError(f"partial_fit: X has {X.shape[1]} features, fitted data has {self._n_features}")
        api, h = self._ensure()
        api.set_feature_range(h, X.shape[1])      # ids are column indices of X: all < X.shape[1]
        if getattr(api, "is_cuda", False) and X.shape[1] <= 0xFFFFFFFF:
            # scipy's arrays go down as they are (float64 values, int32/int64 indices, pageable memory): the library
            # stages them into page-locked memory on several threads, narrows integer counts to uint8/uint16 for the
            # PCIe transfer and overlaps all of it with the upload and K1. Values are only needed by the exact rerank.
            api.fit_csr_host(h, X.indptr, X.indices, None if self.fast else X.data, append)
        else:
            # (CPU back ends of the test-suite; ids above 2^32 go down as 64-bit through the plain entry point)
            data = None if self.fast else self._wire_values(X.data)
            indices = X.indices
            if X.shape[1] > 0xFFFFFFFF:
                indices = indices.astype(np.int64, copy=False)
            elif indices.dtype.itemsize != 4:
                indices = indices.astype(np.uint32)
            api.fit_csr(h, X.indptr, indices, data, append)
        self.number_of_hash_functions = self._created_hashes()    # a fit ends any active-hash prefix
        self._n_features = X.shape[1]
        return self

    def fit(self, X, y=None):
        """Hash all rows of ``X`` (scipy CSR, cells x nbins^2); replaces anything fitted before."""
        return self._fit(X, append=False)

    def partial_fit(self, X, y=None):
        """Append the rows of ``X``; their ids continue after the rows already fitted."""
        if self._h is None or self._api.n_fitted(self._h) == 0:
            return self._fit(X, append=False)
        return self._fit(X, append=True)

    @property
    def n_fitted_(self) -> int:
        return 0 if self._h is None else self._api.n_fitted(self._h)

    # ---- hyperopt fast path ------------------------------------------------------
    def set_number_of_hash_functions(self, n_hashes: int):
        """Switch the query side to the first ``n_hashes`` hash functions WITHOUT re-fitting: h_j depends only on j,
        so that prefix of the fitted signatures is exactly what a fit with ``number_of_hash_functions=n_hashes``
        computes. ``scHicClusterMinHashHyperopt.py:119-148`` re-runs the whole tool per trial with ``-nh`` drawn from a
        range; with this a sweep fits once (at the largest ``-nh``) and keeps the CSR resident on the device.
        ``n_hashes`` must not exceed the value the object was constructed with. The next fit switches back."""
        api, h = self._ensure()
        n_hashes = int(n_hashes)
        if self.shingle and self.shingle_size > 1:
            raise ValueError("set_number_of_hash_functions is not available with block combine (shingle=1)")
        if not 1 <= n_hashes <= self._created_hashes():
            raise ValueError(f"n_hashes must be in [1, {self._created_hashes()}]")
        api.set_active_hashes(h, n_hashes)
        self.number_of_hash_functions = n_hashes
        return self

    def _created_hashes(self) -> int:
        if getattr(self, "_n_hashes_created", None) is None:
            self._n_hashes_created = self.number_of_hash_functions
        return self._n_hashes_created

    def _tables(self) -> int:
        """Columns of the signature matrix: the hash functions, or -- with block combine (``shingle=1``, SURVEY 8(c)
        rule 4; off by default) -- one key per ``shingle_size`` consecutive hash functions."""
        if self.shingle and self.shingle_size > 1:
            return self._created_hashes() // self.shingle_size
        return self.number_of_hash_functions

    # ---- parity hooks --------------------------------------------------------------------------
    def signatures(self) -> np.ndarray:
        api, h = self._ensure()
        return api.signatures(h, self._tables())

    def collision_counts(self, row0=0, row1=None) -> np.ndarray:
        api, h = self._ensure()
        return api.collision_counts(h, row0, row1)

    # ---- queries ---------------------------------------------------------------------------------
    def _fast_flag(self, fast):
        return -1 if fast is None else int(bool(fast))

    def kneighbors(self, X=None, n_neighbors=None, return_distance=True, fast=None):
        """Neighbours of every fitted row (self included, as upstream does for X=None).
        Returns ``(distances, indices)`` of shape [n, k]; missing entries are ``inf`` / ``-1``."""
        if X is not None:
            raise NotImplementedError("kneighbors(X=...) on unfitted rows is not part of the scHiCExplorer path")
        api, h = self._ensure()
        k = int(n_neighbors or self.n_neighbors)
        idx, dist, _ = api.kneighbors(h, k, self._fast_flag(fast))
        return (dist, idx) if return_distance else idx

    def kneighbors_graph(self, X=None, n_neighbors=None, mode="connectivity", fast=None):
        """kNN graph as scipy CSR (n x n); ``mode='distance'`` is what scHicClusterMinHash asks for."""
        if X is not None:
            raise NotImplementedError("kneighbors_graph(X=...) on unfitted rows is not part of the scHiCExplorer path")
        if mode not in ("connectivity", "distance"):
            raise ValueError("mode must be 'connectivity' or 'distance'")
        api, h = self._ensure()
        k = int(n_neighbors or self.n_neighbors)
        n = api.n_fitted(h)
        out = None
        if getattr(api, "is_cuda", False):
            # page-locked result buffers, kept between calls (a fresh pageable array costs its page faults every time)
            ke = min(k, n)
            if getattr(self, "_graph_out", None) is None or len(self._graph_out[0]) < n + 1 or len(self._graph_out[1]) < n * ke:
                self._graph_out = (api.pinned_empty((n + 1,), np.int64), api.pinned_empty((n * ke,), np.int32),
                                   api.pinned_empty((n * ke,), np.float32))
            out = self._graph_out
            k = ke
        indptr, indices, data = api.kneighbors_graph(h, k, self._fast_flag(fast), **({"out": out} if out is not None else {}))
        nnz = len(data)
        values = np.ones(nnz, dtype=np.float64) if mode == "connectivity" else data.astype(np.float64)
        # (copies out of the reused buffers; the arrays come sorted and duplicate-free from K5)
        g = csr_matrix((values, indices.copy(), indptr[:n + 1].copy()), shape=(n, n))
        g.has_sorted_indices = True
        g.has_canonical_format = True
        return g

    def fit_kneighbors_graph(self, X, n_neighbors=None, mode="connectivity", fast=None):
        self.fit(X)
        return self.kneighbors_graph(n_neighbors=n_neighbors, mode=mode, fast=fast)

    def get_params(self, deep=True):
        return {k: getattr(self, k) for k in (
            "n_neighbors", "fast", "number_of_hash_functions", "max_bin_size", "minimal_blocks_in_common",
            "shingle_size", "excess_factor", "number_of_cores", "shingle", "absolute_numbers", "hash_width")}
