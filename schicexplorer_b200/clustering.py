"""Drop-in for ``sparse_neighbors_search.MinHashClustering`` (host side).

Mirrors the call sites ``/root/reference/schicexplorer/scHicClusterMinHash.py:392-393, 405-406,
409-415``: ``fit(X, pSaveMemory=, pPca=, pPcaDimensions=, pUmap=, pUmapDict=)`` builds the kNN
graph with :class:`MinHash` (the GPU hot path), then runs the same host post-processing that the
``--saveMemory`` branch spells out in-tree (:355-390): nan/inf scrub, PCA with ``n-1`` components
cut to ``pPcaDimensions``, optional UMAP, ``clusteringObject.fit``. The PCA runs on the GPU
(:class:`schicexplorer_b200.pca.PCA`, SURVEY 8(f)-2: same numbers as scikit-learn's, only the columns that are
kept are extracted); UMAP / clustering stay on the host (umap-learn, scikit-learn), exactly as in the reference.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import issparse


class MinHashClustering:
    def __init__(self, minHashObject, clusteringObject):
        self._minHashObject = minHashObject
        self._clusteringObject = clusteringObject
        self._precomputed_graph = None

    def fit(self, X, y=None, pSaveMemory=None, pPca=False, pPcaDimensions=None, pUmap=False, pUmapDict=None):
        n = X.shape[0]
        if pSaveMemory is not None and 0 < pSaveMemory < 1:
            # chunked transfer (the reason for -s/--shareOfMatrixToBeTransferred, :94-98)
            step = max(1, int(n * pSaveMemory))
            self._minHashObject.fit(X[0:step])
            for i in range(step, n, step):
                self._minHashObject.partial_fit(X[i:i + step])
        else:
            self._minHashObject.fit(X)
        graph = self._minHashObject.kneighbors_graph(mode="distance")
        graph.data[~np.isfinite(graph.data)] = 0          # :356-357
        self._precomputed_graph = self._post_process(graph, pPca, pPcaDimensions, pUmap, pUmapDict)
        self._fit_clustering(self._precomputed_graph)
        return self

    @staticmethod
    def _post_process(graph, pPca, pPcaDimensions, pUmap, pUmapDict):
        if pPca:
            from .pca import PCA
            n_components = max(1, min(graph.shape) - 1)      # :358-362
            keep = min(pPcaDimensions, graph.shape[0]) if pPcaDimensions else None   # :364-366
            # (non-finite entries are zeroed on the device while the graph is densified)
            graph = PCA(n_components=n_components, keep=keep).fit_transform(graph)
        if pUmap:
            try:
                import umap
            except ImportError as exc:
                raise ImportError("pUmap=True needs the 'umap-learn' package (host-side, not installed here); "
                                  "pass --noUMAP / pUmap=False") from exc
            kw = {}
            if pUmapDict:
                kw = {k[len("umap_"):]: v for k, v in pUmapDict.items() if k.startswith("umap_") and k != "umap_random"}
                kw["random_state"] = pUmapDict.get("umap_random", 42)
            dense = np.asarray(graph.todense()) if issparse(graph) else graph
            graph = umap.UMAP(**kw).fit_transform(dense)     # :369-383
        if not issparse(graph):
            graph = np.nan_to_num(graph)
            graph[np.isinf(graph)] = 0
        return graph

    def _fit_clustering(self, graph):
        try:
            self._clusteringObject.fit(graph)
        except Exception:
            self._clusteringObject.fit(np.asarray(graph.todense()))   # :387-390

    def fit_predict(self, X, y=None, **kw):
        self.fit(X, **kw)
        return self.predict(self._precomputed_graph)

    def predict(self, X=None, y=None, pPca=None, pPcaDimensions=None):
        """Cluster labels of the fitted cells (:415). ``X`` is the precomputed graph the caller holds."""
        obj = self._clusteringObject
        if hasattr(obj, "labels_"):
            return np.asarray(obj.labels_)
        if X is None:
            X = self._precomputed_graph
        if hasattr(obj, "fit_predict"):
            return np.asarray(obj.fit_predict(X))
        return np.asarray(obj.predict(X))
