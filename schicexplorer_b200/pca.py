"""Dense-graph PCA on the GPU: the post-step of scHicClusterMinHash / scHicCluster.

Boundary mirrored: ``sklearn.decomposition.PCA`` as the reference uses it --
``PCA(n_components=min(shape) - 1).fit_transform(graph.todense())`` followed by ``[:, :dimensionsPCA]``
(``/root/reference/schicexplorer/scHicClusterMinHash.py:358-366``, ``scHicCluster.py:183-188``). Same constructor
argument, ``fit_transform`` / ``fit``, ``components_``, ``singular_values_``, ``mean_``, ``explained_variance_``. ``keep`` (the only addition) tells the device how many leading columns the caller will
actually use, so that only those eigenpairs are extracted; the numbers are those of the full decomposition.
Every number comes from ``csrc/pca_dense.cu`` through ``include/schic_pca.h``; no CPU fallback."""
from __future__ import annotations

import numpy as np
from scipy.sparse import issparse

from . import _capi


class PCA:
    _api_factory = staticmethod(_capi.cuda_api)

    def __init__(self, n_components=None, keep=None, device=-1, **kwargs):
        # svd_solver / whiten / random_state ... of sklearn's constructor: this class is the full-SVD, non-whitening
        # PCA the reference asks for; other settings are refused rather than silently ignored
        unsupported = {k: v for k, v in kwargs.items() if not (k == "svd_solver" and v in ("auto", "full")) and
                       not (k == "whiten" and not v) and k not in ("random_state", "copy")}
        if unsupported:
            raise ValueError(f"unsupported PCA arguments: {sorted(unsupported)}")
        self.n_components = n_components
        self.keep = keep
        self.device = int(device)

    def _columns(self, shape) -> int:
        full = min(shape) if self.n_components is None else int(self.n_components)
        if not 1 <= full <= min(shape):
            raise ValueError(f"n_components={full} must be between 1 and min(n_samples, n_features)={min(shape)}")
        return full if self.keep is None else max(1, min(full, int(self.keep)))

    def fit_transform(self, X, y=None):
        """Embedding U * S, [n_samples, min(n_components, keep)]. X: dense array or scipy sparse (densified on the device)."""
        if not issparse(X):
            X = np.asarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("X must be two-dimensional")
        m = self._columns(X.shape)
        api = self._api_factory()
        out, sv, comp, mean = api.pca_fit_transform(X, m, self.device, want_components=True, return_mean=True)
        n = X.shape[0]
        self.components_ = comp
        self.singular_values_ = sv
        self.explained_variance_ = sv ** 2 / max(n - 1, 1)
        self.mean_ = mean
        self.n_components_ = m
        return out

    def fit(self, X, y=None):
        self.fit_transform(X)
        return self
