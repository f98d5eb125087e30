"""Exact k-nearest-neighbour graph on sparse rows: the ``scHicCluster -drm knn`` step on the GPU.

Boundary mirrored: ``/root/reference/schicexplorer/scHicCluster.py:179-180`` --

    nbrs = NearestNeighbors(n_neighbors=k, algorithm='ball_tree', n_jobs=threads).fit(neighborhood_matrix)
    neighborhood_matrix = nbrs.kneighbors_graph(mode='distance')

(scikit-learn; on a sparse matrix this is brute-force euclidean kNN, and with ``X=None`` the query row itself is left
out of its own neighbour list). This class offers that subset of ``sklearn.neighbors.NearestNeighbors`` with the same
names and argument meaning. Every number comes from the K4 kernels (``csrc/k4_rerank.cu``) through
``schic_mh_kneighbors_exact*`` of ``include/schic_mh.h``; no CPU fallback.
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy.sparse import csr_matrix

from . import _capi
from .minhash import MinHash


class NearestNeighbors:
    """``sklearn.neighbors.NearestNeighbors`` as scHicCluster uses it (euclidean, queries = the fitted rows)."""

    _api_factory = staticmethod(_capi.cuda_api)

    def __init__(self, n_neighbors=5, algorithm="auto", n_jobs=None, metric="minkowski", p=2, device=-1, **kwargs):
        if metric not in ("minkowski", "euclidean", "l2") or (metric == "minkowski" and p != 2):
            raise ValueError("only the euclidean metric is supported")
        if kwargs:
            warnings.warn(f"NearestNeighbors: ignoring unsupported keyword arguments {sorted(kwargs)}")
        self.n_neighbors = int(n_neighbors)
        self.algorithm = algorithm      # accepted for signature compatibility: the search is always brute force
        self.n_jobs = n_jobs
        self.device = int(device)
        self._api = None
        self._h = None
        self.n_samples_fit_ = 0

    def _ensure(self):
        if self._h is None:
            self._api = self._api_factory()
            # the MinHash side of the handle is idle here: one hash function keeps its cost at nothing
            self._h = self._api.create(n_neighbors=self.n_neighbors, number_of_hash_functions=1, fast=False,
                                       minimal_blocks_in_common=1, excess_factor=1, max_bin_size=1 << 40,
                                       device=self.device)
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

    def fit(self, X, y=None):
        X = MinHash._as_csr(X)
        api, h = self._ensure()
        indices = X.indices
        if X.shape[1] > 0xFFFFFFFF:
            raise ValueError("exact kNN supports feature ids below 2^32")
        if indices.dtype.itemsize != 4:
            indices = indices.astype(np.uint32)
        api.set_feature_range(h, X.shape[1])
        api.fit_csr(h, X.indptr, indices, MinHash._wire_values(X.data), False)
        self.n_samples_fit_ = X.shape[0]
        return self

    def _k(self, n_neighbors, include_self):
        k = int(n_neighbors or self.n_neighbors)
        limit = self.n_samples_fit_ - (0 if include_self else 1)
        if k < 1 or k > limit:
            raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {k}, n_samples_fit = "
                             f"{self.n_samples_fit_}" + ("" if include_self else " (the query itself is excluded)"))
        return k

    def kneighbors(self, X=None, n_neighbors=None, return_distance=True):
        """Neighbours of every fitted row, the row itself excluded (sklearn's ``X=None`` convention).
        Returns ``(distances, indices)`` of shape [n, k], float64 / int64 as sklearn does."""
        if X is not None:
            raise NotImplementedError("kneighbors(X=...) on unfitted rows is not part of the scHiCExplorer path")
        api, h = self._ensure()
        idx, dist, _ = api.kneighbors_exact(h, self._k(n_neighbors, False), include_self=False)
        idx = idx.astype(np.int64)
        return (dist.astype(np.float64), idx) if return_distance else idx

    def kneighbors_graph(self, X=None, n_neighbors=None, mode="connectivity"):
        """kNN graph as scipy CSR (n x n), exactly k entries per row; ``mode='distance'`` is what scHicCluster asks for."""
        if X is not None:
            raise NotImplementedError("kneighbors_graph(X=...) on unfitted rows is not part of the scHiCExplorer path")
        if mode not in ("connectivity", "distance"):
            raise ValueError('Unsupported mode, must be one of "connectivity", or "distance" but got "%s" instead' % mode)
        api, h = self._ensure()
        n = self.n_samples_fit_
        indptr, indices, data = api.kneighbors_exact_graph(h, self._k(n_neighbors, False), include_self=False)
        data = np.ones(len(data)) if mode == "connectivity" else data.astype(np.float64)
        return csr_matrix((data, indices, indptr), shape=(n, n))
