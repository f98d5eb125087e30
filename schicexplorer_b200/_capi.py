"""ctypes binding of the C ABI declared in ``include/schic_mh.h`` / ``include/schic_mh_cuda.h``.

The product loads exactly one library through :func:`cuda_api`:
``schicexplorer_b200/csrc/libschic_mh.so`` (hand-written sm_100a kernels). There is no CPU
fallback: if the library is missing, or no GPU is usable, the call raises. The class
:class:`CApi` itself is path-agnostic so that the tests can bind the oracle's ``.so`` (which
exports the same symbols) as the checker; nothing in this package references ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCHIC_CUDA_LIB: experiment builds of the same library (tools/); the product always loads the in-tree default
CUDA_LIB_PATH = os.environ.get("SCHIC_CUDA_LIB") or os.path.join(_HERE, "csrc", "libschic_mh.so")


class SchicError(RuntimeError):
    """A C-ABI call returned a negative status (message from ``schic_mh_last_error``)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[schic status {status}] {message}")
        self.status = status


class Params(C.Structure):
    # field order mirrors struct schic_mh_params (include/schic_mh.h)
    _fields_ = [
        ("struct_size", C.c_int32),
        ("n_neighbors", C.c_int32),
        ("number_of_hash_functions", C.c_int32),
        ("hash_width", C.c_int32),
        ("fast", C.c_int32),
        ("absolute_numbers", C.c_int32),
        ("minimal_blocks_in_common", C.c_int32),
        ("excess_factor", C.c_int32),
        ("max_bin_size", C.c_int64),
        ("shingle_size", C.c_int32),
        ("shingle", C.c_int32),
        ("number_of_cores", C.c_int32),
        ("device", C.c_int32),
    ]


# symbols every back end must export (include/schic_mh.h)
BASE_SYMBOLS = (
    "schic_mh_create", "schic_mh_destroy", "schic_mh_last_error", "schic_mh_backend",
    "schic_mh_fit_csr", "schic_mh_n_fitted", "schic_mh_signatures", "schic_mh_collision_counts",
    "schic_mh_kneighbors", "schic_mh_kneighbors_graph", "schic_mh_kneighbors_exact", "schic_mh_kneighbors_exact_graph",
    "schic_mh_set_active_hashes",
)
# symbols only the CUDA library exports (include/schic_mh_cuda.h)
CUDA_SYMBOLS = (
    "schic_mh_fit_csr_async", "schic_mh_fit_csr_counts", "schic_mh_set_feature_range", "schic_mh_fit_csr_device", "schic_mh_kneighbors_device", "schic_mh_signatures_device", "schic_mh_csr_device",
    "schic_mh_set_database_device", "schic_mh_set_database_ready_event", "schic_mh_rerank_windows",
    "schic_mh_rerank_prep_rows", "schic_mh_set_database_prep_device", "schic_mh_set_stream", "schic_mh_stage_ms",
    "schic_mh_launch_count", "schic_mh_synchronize", "schic_int32_peak", "schic_device_count",
    "schic_k1_variant_ms", "schic_pinned_alloc", "schic_pinned_free", "schic_mh_kneighbors_exact_device",
    "schic_mh_fit_csr_device_at", "schic_mh_kneighbors_graph_device", "schic_mh_rerank_kernel_name", "schic_mh_k1_table_state", "schic_mh_fit_csr_host", "schic_mh_collision_block_device", "schic_mh_set_counts_device",
)

# the dense-graph PCA post-step, CUDA library only (include/schic_pca.h)
PCA_SYMBOLS = ("schic_pca_fit_transform", "schic_pca_fit_transform_csr", "schic_pca_stage_ms")

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(_vp)
    return _vp(int(a))  # raw (device) address


class CApi:
    """One loaded back end. All methods raise :class:`SchicError` on a non-zero status."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.schic_mh_last_error.restype = C.c_char_p
        L.schic_mh_backend.restype = C.c_char_p
        L.schic_mh_create.argtypes = [C.POINTER(Params), C.POINTER(_vp)]
        L.schic_mh_destroy.argtypes = [_vp]
        L.schic_mh_fit_csr.argtypes = [_vp, _i64, _vp, _vp, _i32, _vp, _i32]
        L.schic_mh_n_fitted.argtypes = [_vp, C.POINTER(_i64)]
        L.schic_mh_signatures.argtypes = [_vp, _vp]
        L.schic_mh_collision_counts.argtypes = [_vp, _i64, _i64, _vp]
        L.schic_mh_kneighbors.argtypes = [_vp, _i32, _i32, _vp, _vp, _vp]
        L.schic_mh_kneighbors_graph.argtypes = [_vp, _i32, _i32, _vp, _vp, _vp]
        L.schic_mh_kneighbors_exact.argtypes = [_vp, _i32, _i32, _vp, _vp, _vp]
        L.schic_mh_kneighbors_exact_graph.argtypes = [_vp, _i32, _i32, _vp, _vp, _vp]
        L.schic_mh_set_active_hashes.argtypes = [_vp, _i32]
        self.backend = L.schic_mh_backend().decode()
        self.is_cuda = self.backend == "cuda"
        if self.is_cuda:
            L.schic_mh_fit_csr_async.argtypes = [_vp, _i64, _vp, _vp, _i32, _vp, _i32]
            L.schic_mh_set_feature_range.argtypes = [_vp, C.c_uint64]
            L.schic_mh_kneighbors_exact_device.argtypes = [_vp, _i32, _i32, _vp, _vp, _vp]
            L.schic_pca_fit_transform.argtypes = [_vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _vp]
            L.schic_pca_fit_transform_csr.argtypes = [_i64, _i64, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]
            L.schic_pca_stage_ms.argtypes = [_vp, _i32]
            L.schic_mh_set_database_ready_event.argtypes = [_vp, _vp]
            L.schic_mh_rerank_windows.argtypes = [_vp, C.POINTER(_i32)]
            L.schic_mh_rerank_prep_rows.argtypes = [_vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]
            L.schic_mh_set_database_prep_device.argtypes = [_vp, _vp, _vp, _i32, _vp, _vp, _i64]
            L.schic_mh_rerank_kernel_name.argtypes = [_vp, C.c_char_p, _i32]
            L.schic_mh_fit_csr_counts.argtypes = [_vp, _i64, _vp, _vp, _i32, _vp, _i32, _i32, _i32]
            L.schic_mh_fit_csr_device.argtypes = [_vp, _i64, _i64, _vp, _vp, _vp, _i32]
            L.schic_mh_kneighbors_device.argtypes = [_vp, _i32, _i32, _vp, _vp, _vp]
            L.schic_mh_collision_block_device.argtypes = [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _i64]
            L.schic_mh_set_counts_device.argtypes = [_vp, _vp, _i64]
            L.schic_mh_fit_csr_host.argtypes = [_vp, _i64, _vp, _i32, _vp, _i32, _vp, _i32, _i32, _i32]
            L.schic_mh_fit_csr_device_at.argtypes = [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32]
            L.schic_mh_kneighbors_graph_device.argtypes = [_vp, _i32, _i32, _vp, _vp, _vp]
            L.schic_mh_signatures_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_i64)]
            L.schic_mh_csr_device.argtypes = [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64)]
            L.schic_mh_set_database_device.argtypes = [_vp, _i64, _i64, _vp, _vp, _vp, _vp]
            L.schic_mh_set_stream.argtypes = [_vp, _vp]
            L.schic_mh_stage_ms.argtypes = [_vp, _vp, _i32]
            L.schic_mh_launch_count.argtypes = [_vp, C.POINTER(_i64)]
            L.schic_mh_k1_table_state.argtypes = [_vp, C.POINTER(_i32)]
            L.schic_mh_synchronize.argtypes = [_vp]
            L.schic_int32_peak.argtypes = [_i32, _i32, _vp]
            L.schic_k1_variant_ms.argtypes = [_i32, _i64, _i64, _vp, _vp, _i32, _vp, _i32, C.POINTER(C.c_float)]
            L.schic_pinned_alloc.argtypes = [_i64, C.POINTER(_vp)]
            L.schic_pinned_free.argtypes = [_vp]
            L.schic_device_count.argtypes = [C.POINTER(_i32)]

    # -- helpers ---------------------------------------------------------------------------
    def check(self, status: int):
        if status != 0:
            raise SchicError(status, self.lib.schic_mh_last_error().decode(errors="replace"))

    def has_symbols(self, names) -> list:
        missing = []
        for n in names:
            try:
                getattr(self.lib, n)
            except AttributeError:
                missing.append(n)
        return missing

    # -- include/schic_mh.h ----------------------------------------------------------------
    def create(self, **kw) -> int:
        p = Params()
        p.struct_size = C.sizeof(Params)
        p.n_neighbors = int(kw.get("n_neighbors", 5))
        p.number_of_hash_functions = int(kw.get("number_of_hash_functions", 400))
        p.hash_width = int(kw.get("hash_width", 32))
        p.fast = int(bool(kw.get("fast", False)))
        p.absolute_numbers = int(bool(kw.get("absolute_numbers", False)))
        p.minimal_blocks_in_common = int(kw.get("minimal_blocks_in_common", 1))
        p.excess_factor = int(kw.get("excess_factor", 1))
        p.max_bin_size = int(kw.get("max_bin_size", 50))
        p.shingle_size = int(kw.get("shingle_size", 4))
        p.shingle = int(kw.get("shingle", 0))
        p.number_of_cores = int(kw.get("number_of_cores", 0) or 0)
        p.device = int(kw.get("device", -1))
        h = _vp()
        self.check(self.lib.schic_mh_create(C.byref(p), C.byref(h)))
        return h.value

    def destroy(self, h):
        if h:
            self.lib.schic_mh_destroy(_vp(h))

    def fit_csr(self, h, indptr, indices, data, append: bool, async_upload: bool = False):
        """``async_upload`` (CUDA back end only): return once the uploads and K1 are queued; the caller keeps the
        three arrays alive until the next synchronising call (include/schic_mh_cuda.h). The arrays must then
        already have the wire dtypes (int64 / uint32|uint64 / float32, C contiguous) so no temporary is made."""
        if async_upload:
            ok = (indptr.dtype == np.int64 and indptr.flags.c_contiguous and indices.flags.c_contiguous and
                  indices.dtype.itemsize in (4, 8) and
                  (data is None or (data.dtype in (np.float32, np.uint16, np.uint8) and data.flags.c_contiguous)))
            if not ok:
                raise ValueError("async_upload needs int64 indptr, 32/64-bit indices and float32 data, C contiguous")
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        if indices.dtype.itemsize == 8:
            indices = np.ascontiguousarray(indices).view(np.uint64)
            bits = 64
        else:
            indices = np.ascontiguousarray(indices).view(np.uint32)
            bits = 32
        if data is not None and data.dtype in (np.uint8, np.uint16) and self.is_cuda:
            # integer counts cross PCIe narrow and are widened on the device (include/schic_mh_cuda.h)
            data = np.ascontiguousarray(data)
            self.check(self.lib.schic_mh_fit_csr_counts(_vp(h), len(indptr) - 1, _ptr(indptr), _ptr(indices), bits,
                                                        _ptr(data), data.dtype.itemsize * 8, int(append), int(async_upload)))
            return
        if data is not None:
            data = np.ascontiguousarray(data, dtype=np.float32)
        fn = self.lib.schic_mh_fit_csr_async if async_upload else self.lib.schic_mh_fit_csr
        self.check(fn(_vp(h), len(indptr) - 1, _ptr(indptr), _ptr(indices), bits, _ptr(data), int(append)))

    _HOST_KINDS = {np.dtype(np.float32): 0, np.dtype(np.uint16): 1, np.dtype(np.uint8): 2, np.dtype(np.float64): 3}

    def fit_csr_host(self, h, indptr, indices, data, append: bool, async_upload: bool = True):
        """CUDA back end: fit from scipy's own arrays (int32/int64 indptr and indices, float64/float32/uint8/uint16 data,
        pageable memory) without a Python-side copy; staging, narrowing and upload happen inside the library on several
        threads (include/schic_mh_cuda.h). Other dtypes are converted here first."""
        if indptr.dtype.itemsize not in (4, 8) or indptr.dtype.kind not in "iu":
            indptr = indptr.astype(np.int64)
        if indices.dtype.itemsize not in (4, 8) or indices.dtype.kind not in "iu":
            indices = indices.astype(np.int64)
        indptr, indices = np.ascontiguousarray(indptr), np.ascontiguousarray(indices)
        kind = -1
        if data is not None:
            if data.dtype not in self._HOST_KINDS:
                data = data.astype(np.float64)
            data = np.ascontiguousarray(data)
            kind = self._HOST_KINDS[data.dtype]
        self.check(self.lib.schic_mh_fit_csr_host(_vp(h), len(indptr) - 1, _ptr(indptr), indptr.dtype.itemsize * 8,
                                                  _ptr(indices), indices.dtype.itemsize * 8, _ptr(data), kind, int(append),
                                                  int(async_upload)))

    def n_fitted(self, h) -> int:
        n = _i64()
        self.check(self.lib.schic_mh_n_fitted(_vp(h), C.byref(n)))
        return n.value

    def signatures(self, h, n_hash: int) -> np.ndarray:
        out = np.empty((self.n_fitted(h), n_hash), dtype=np.uint32)
        self.check(self.lib.schic_mh_signatures(_vp(h), _ptr(out)))
        return out

    def collision_counts(self, h, row0=0, row1=None) -> np.ndarray:
        n = self.n_fitted(h)
        row1 = n if row1 is None else row1
        out = np.empty((row1 - row0, n), dtype=np.uint16)
        self.check(self.lib.schic_mh_collision_counts(_vp(h), row0, row1, _ptr(out)))
        return out

    def kneighbors(self, h, k: int, fast: int = -1, n_rows=None):
        n = self.n_fitted(h) if n_rows is None else n_rows
        idx = np.empty((n, k), dtype=np.int32)
        dist = np.empty((n, k), dtype=np.float32)
        nv = np.empty(n, dtype=np.int32)
        self.check(self.lib.schic_mh_kneighbors(_vp(h), k, fast, _ptr(idx), _ptr(dist), _ptr(nv)))
        return idx, dist, nv

    def kneighbors_graph(self, h, k: int, fast: int = -1, n_rows=None, out=None):
        """``out``: optional (indptr int64 [n+1], indices int32 [n*k], data float32 [n*k]) host arrays to fill (e.g.
        page-locked ones that are reused between calls)."""
        n = self.n_fitted(h) if n_rows is None else n_rows
        if out is not None:
            indptr, indices, data = out
            assert indptr.dtype == np.int64 and indices.dtype == np.int32 and data.dtype == np.float32
            assert len(indptr) >= n + 1 and len(indices) >= n * k and len(data) >= n * k
        else:
            indptr = np.empty(n + 1, dtype=np.int64)
            indices = np.empty(n * k, dtype=np.int32)
            data = np.empty(n * k, dtype=np.float32)
        self.check(self.lib.schic_mh_kneighbors_graph(_vp(h), k, fast, _ptr(indptr), _ptr(indices), _ptr(data)))
        nnz = int(indptr[-1])
        return indptr, indices[:nnz], data[:nnz]

    def kneighbors_exact(self, h, k: int, include_self: bool = False):
        """Brute-force euclidean kNN over the fitted rows (scHicCluster -drm knn, scHicCluster.py:179-180)."""
        n = self.n_fitted(h)
        idx = np.empty((n, k), dtype=np.int32)
        dist = np.empty((n, k), dtype=np.float32)
        nv = np.empty(n, dtype=np.int32)
        self.check(self.lib.schic_mh_kneighbors_exact(_vp(h), k, int(include_self), _ptr(idx), _ptr(dist), _ptr(nv)))
        return idx, dist, nv

    def kneighbors_exact_graph(self, h, k: int, include_self: bool = False):
        n = self.n_fitted(h)
        indptr = np.empty(n + 1, dtype=np.int64)
        indices = np.empty(n * k, dtype=np.int32)
        data = np.empty(n * k, dtype=np.float32)
        self.check(self.lib.schic_mh_kneighbors_exact_graph(_vp(h), k, int(include_self), _ptr(indptr), _ptr(indices),
                                                            _ptr(data)))
        nnz = int(indptr[-1])
        return indptr, indices[:nnz], data[:nnz]

    def set_active_hashes(self, h, n_hashes: int):
        """Query side uses the first n_hashes hash functions of the fitted signatures (hyperopt sweeps over -nh)."""
        self.check(self.lib.schic_mh_set_active_hashes(_vp(h), int(n_hashes)))

    # -- include/schic_mh_cuda.h (device-pointer entry points) -------------------------------
    def kneighbors_exact_device(self, h, k, include_self, d_idx, d_dist, d_nvalid):
        self.check(self.lib.schic_mh_kneighbors_exact_device(_vp(h), k, int(include_self), _ptr(d_idx), _ptr(d_dist),
                                                             _ptr(d_nvalid)))

    def fit_csr_device(self, h, n_rows, nnz, d_indptr, d_indices, d_data, append: bool, indptr0=None):
        """``indptr0``: the value of d_indptr[0] when the caller knows it (0 for a stand-alone shard) -- the call then
        enqueues without reading anything back."""
        if indptr0 is not None and self.is_cuda:
            self.check(self.lib.schic_mh_fit_csr_device_at(_vp(h), n_rows, nnz, int(indptr0), _ptr(d_indptr),
                                                           _ptr(d_indices), _ptr(d_data), int(append)))
            return
        self.check(self.lib.schic_mh_fit_csr_device(_vp(h), n_rows, nnz, _ptr(d_indptr), _ptr(d_indices),
                                                    _ptr(d_data), int(append)))

    def kneighbors_graph_device(self, h, k, fast, d_indptr, d_indices, d_data):
        self.check(self.lib.schic_mh_kneighbors_graph_device(_vp(h), k, fast, _ptr(d_indptr), _ptr(d_indices),
                                                             _ptr(d_data)))

    def kneighbors_device(self, h, k, fast, d_idx, d_dist, d_nvalid):
        self.check(self.lib.schic_mh_kneighbors_device(_vp(h), k, fast, _ptr(d_idx), _ptr(d_dist), _ptr(d_nvalid)))

    def signatures_device(self, h):
        p, n = _vp(), _i64()
        self.check(self.lib.schic_mh_signatures_device(_vp(h), C.byref(p), C.byref(n)))
        return p.value, n.value

    def csr_device(self, h):
        a, b, c, n = _vp(), _vp(), _vp(), _i64()
        self.check(self.lib.schic_mh_csr_device(_vp(h), C.byref(a), C.byref(b), C.byref(c), C.byref(n)))
        return a.value, b.value, c.value, n.value

    # -- include/schic_pca.h ---------------------------------------------------------------------
    def pca_fit_transform(self, X, n_components: int, device: int = -1, want_components: bool = False,
                          return_mean: bool = False):
        """PCA(...).fit_transform(X)[:, :n_components] on the device. X: dense float64 [n, c] or a scipy CSR. A CSR goes
        down in the graph entry points' layout (int64 / int32 / float32) when its values are exactly representable in
        float32 (every graph this package produces is); otherwise it is densified on the host and takes the float64
        dense entry point, so nothing is rounded silently. Returns (embedding, singular values, components or None),
        plus the column means with ``return_mean`` (``last_pca_mean`` on the shared object is kept for older callers
        only -- it is not safe with two PCA objects or threads)."""
        from scipy.sparse import issparse
        if issparse(X):
            X = X.tocsr()
            if X.data.dtype != np.float32 and not np.array_equal(X.data.astype(np.float32).astype(X.data.dtype), X.data):
                X = np.asarray(X.todense(), dtype=np.float64)
        n_rows, n_cols = X.shape
        m = int(n_components)
        out = np.empty((n_rows, m), dtype=np.float64)
        sv = np.empty(m, dtype=np.float64)
        comp = np.empty((m, n_cols), dtype=np.float64) if want_components else None
        mean = np.empty(n_cols, dtype=np.float64) if want_components else None
        self.last_pca_mean = mean
        if issparse(X):
            X = X.tocsr()
            indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
            indices = np.ascontiguousarray(X.indices, dtype=np.int32)
            data = np.ascontiguousarray(X.data, dtype=np.float32)
            self.check(self.lib.schic_pca_fit_transform_csr(n_rows, n_cols, _ptr(indptr), _ptr(indices), _ptr(data), m,
                                                            device, _ptr(out), _ptr(sv), _ptr(comp), _ptr(mean)))
        else:
            Xd = np.ascontiguousarray(X, dtype=np.float64)
            self.check(self.lib.schic_pca_fit_transform(_ptr(Xd), n_rows, n_cols, m, device, _ptr(out), _ptr(sv), _ptr(comp),
                                                        _ptr(mean)))
        return (out, sv, comp, mean) if return_mean else (out, sv, comp)

    def pca_stage_ms(self) -> dict:
        out = np.zeros(7, dtype=np.float32)
        self.check(self.lib.schic_pca_stage_ms(_ptr(out), 7))
        names = ("upload_centre", "covariance", "tridiagonalise", "eigenvalues", "eigenvectors", "project", "total")
        return dict(zip(names, out.tolist()))

    def set_database_device(self, h, n_total, row0, d_sig, d_indptr=None, d_indices=None, d_data=None):
        self.check(self.lib.schic_mh_set_database_device(_vp(h), n_total, row0, _ptr(d_sig), _ptr(d_indptr),
                                                         _ptr(d_indices), _ptr(d_data)))

    def set_database_ready_event(self, h, cuda_event: int):
        """The CSR arrays given to set_database_device are complete once this cudaEvent_t has fired; the library waits
        for it only before the exact rerank (CUDA back end)."""
        self.check(self.lib.schic_mh_set_database_ready_event(_vp(h), _vp(cuda_event)))

    def rerank_windows(self, h) -> int:
        n = _i32()
        self.check(self.lib.schic_mh_rerank_windows(_vp(h), C.byref(n)))
        return n.value

    def rerank_prep_rows(self, h, n_rows, d_indptr, d_indices, d_data, n_windows, d_norm2_out, d_dir_out, d_flags_out,
                         d_packed_out=None):
        """Row norms, window directory, flags[4] and (optionally) the packed stream of the given CSR rows."""
        self.check(self.lib.schic_mh_rerank_prep_rows(_vp(h), n_rows, _ptr(d_indptr), _ptr(d_indices), _ptr(d_data),
                                                      n_windows, _ptr(d_norm2_out), _ptr(d_dir_out), _ptr(d_packed_out),
                                                      _ptr(d_flags_out)))

    def set_database_prep_device(self, h, d_norm2_all, d_dir_all, n_windows, d_packed_all=None, d_flags_all=None,
                                 nnz_total=0):
        self.check(self.lib.schic_mh_set_database_prep_device(_vp(h), _ptr(d_norm2_all), _ptr(d_dir_all), n_windows,
                                                              _ptr(d_packed_all), _ptr(d_flags_all), int(nnz_total)))

    def collision_block_device(self, h, q_row0, q_rows, d_row0, d_rows, d_cnt, ld, d_cnt_t=None, ld_t=0):
        """Counts of database rows [q_row0, +q_rows) x [d_row0, +d_rows) into d_cnt (row stride ld), optionally also
        transposed into d_cnt_t (row stride ld_t). Device pointers; asynchronous."""
        self.check(self.lib.schic_mh_collision_block_device(_vp(h), q_row0, q_rows, d_row0, d_rows, _ptr(d_cnt), ld,
                                                            _ptr(d_cnt_t), ld_t))

    def set_counts_device(self, h, d_counts, ld):
        """The next kneighbors* call uses this [local rows, ld] uint16 matrix as its collision counts."""
        self.check(self.lib.schic_mh_set_counts_device(_vp(h), _ptr(d_counts), ld))

    def rerank_kernel_name(self, h) -> str:
        """Which kernel did the work of the last exact rerank (counts kernel or a generic float kernel)."""
        buf = C.create_string_buffer(64)
        self.check(self.lib.schic_mh_rerank_kernel_name(_vp(h), buf, 64))
        return buf.value.decode()

    def set_feature_range(self, h, n_features: int):
        """Hint (CUDA back end): every feature id is < n_features (the CSR's column count); lets fit / kneighbors
        enqueue without reading the largest id back. No-op on back ends without the entry point."""
        fn = getattr(self.lib, "schic_mh_set_feature_range", None)
        if fn is not None and self.backend == "cuda":
            self.check(fn(_vp(h), int(n_features)))

    def set_stream(self, h, stream: int):
        self.check(self.lib.schic_mh_set_stream(_vp(h), _vp(stream)))

    def stage_ms(self, h) -> dict:
        out = np.zeros(9, dtype=np.float32)
        self.check(self.lib.schic_mh_stage_ms(_vp(h), _ptr(out), 9))
        names = ("h2d", "signatures", "collisions", "select", "rerank", "graph", "d2h", "rerank_prep", "total")
        return dict(zip(names, out.tolist()))

    K1_TABLE_STATES = ("none", "built from the batch's ids", "whole-range table built", "whole-range table re-used")

    def k1_table_state(self, h) -> int:
        """0 no table, 1 built from the batch's ids, 2 whole-range table built, 3 cached whole-range table re-used."""
        n = _i32()
        self.check(self.lib.schic_mh_k1_table_state(_vp(h), C.byref(n)))
        return n.value

    def launch_count(self, h) -> int:
        n = _i64()
        self.check(self.lib.schic_mh_launch_count(_vp(h), C.byref(n)))
        return n.value

    def synchronize(self, h):
        self.check(self.lib.schic_mh_synchronize(_vp(h)))

    def int32_peak(self, device: int = -1, iters: int = 4096) -> dict:
        """INT32 issue-rate micro-benchmark (Tint-op/s per SASS opcode) -- the roofline denominator."""
        out = np.zeros(16, dtype=np.float64)
        self.check(self.lib.schic_int32_peak(device, iters, _ptr(out)))
        names = ("lop3", "imad", "mix_lop3_imad", "sm_mhz", "shf", "imad_hi", "vimnmx", "vimnmx3", "iadd", "imad_wide",
                 "hset2", "hadd2", "mix_hset2_hadd2", "isetp_fadd_pair")
        return dict(zip(names, out.tolist()))

    def k1_variant_ms(self, variant, n_rows, nnz, d_indptr, d_indices, n_hash, d_sig, reps=3) -> float:
        ms = C.c_float()
        self.check(self.lib.schic_k1_variant_ms(variant, n_rows, nnz, _ptr(d_indptr), _ptr(d_indices), n_hash,
                                                _ptr(d_sig), reps, C.byref(ms)))
        return ms.value

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array over page-locked host memory (freed when the array is collected)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = _vp()
        self.check(self.lib.schic_pinned_alloc(n, C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        lib, addr = self.lib, p.value
        import weakref
        weakref.finalize(buf, lambda: lib.schic_pinned_free(_vp(addr)))
        return arr

    def device_count(self) -> int:
        n = _i32()
        self.check(self.lib.schic_device_count(C.byref(n)))
        return n.value


_cuda_api = None
_lock = threading.Lock()


def cuda_api() -> CApi:
    """The product back end. Raises if the CUDA library is not built (no fallback)."""
    global _cuda_api
    with _lock:
        if _cuda_api is None:
            api = CApi(CUDA_LIB_PATH)
            if not api.is_cuda:
                raise RuntimeError(f"{CUDA_LIB_PATH} reports backend {api.backend!r}, expected 'cuda'")
            _cuda_api = api
    return _cuda_api
