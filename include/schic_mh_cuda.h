/*
 * schic_mh_cuda.h -- entry points only the CUDA library (libschic_mh.so) exports, on top of
 * include/schic_mh.h. They take DEVICE pointers (plain addresses, no torch types) so that
 *   - bench.py can time the path with inputs already resident in HBM,
 *   - the multi-GPU host layer (one process per GPU) can all-gather signatures / CSR shards with
 *     NCCL through torch.distributed and hand the gathered buffers back to the kernels.
 * The reference has no counterpart: scHiCExplorer never enables upstream's GPU flags
 * (no gpu_hashing kwarg at scHicClusterMinHash.py:347-348, 402-404); these calls serve the same
 * MinHash.fit / kneighbors sites as schic_mh.h.
 */
#ifndef SCHIC_MH_CUDA_H_
#define SCHIC_MH_CUDA_H_

#include "schic_mh.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Number of visible CUDA devices (0 and SCHIC_OK when there is no driver/GPU). */
int schic_device_count(int32_t* n_out);

/* schic_mh_fit_csr (host CSR in) that returns as soon as the uploads and K1 are queued instead of when the
 * last byte has crossed PCIe: the value array is only read by the exact rerank, so collision counting can
 * start while it is still in flight. The caller keeps indptr / indices / data alive and unchanged until
 * schic_mh_synchronize(), a kneighbors call that returns host results, or the next fit on the handle. */
int schic_mh_fit_csr_async(schic_mh_handle* h, int64_t n_rows, const int64_t* indptr, const void* indices,
                           int32_t index_bits, const float* data, int32_t append);

/* MinHash.fit straight from the arrays a scipy.sparse.csr_matrix holds (the matrix utilities.py:131 builds: float64 data,
 * int32 or int64 indptr / indices, pageable memory): indptr_bits / index_bits = 32 | 64, data_kind = 0 float32, 1 uint16,
 * 2 uint8, 3 float64 (data may be NULL when fast=1). A pool of host threads (params.number_of_cores, default: all) stages
 * the arrays chunk by chunk into page-locked buffers owned by the handle -- ids as 32-bit words, values narrowed to
 * uint8 / uint16 counts wherever a chunk holds only such integers (Hi-C counts do), else float32 -- while the calling
 * thread uploads the chunks that are ready and queues K1 behind them: conversion, PCIe and hashing overlap, nothing is
 * validated or copied by the Python side. async_upload != 0: return once everything is queued; the caller's arrays are
 * no longer needed when the call returns (the staging buffers are the source of the copies). Ids >= 2^32 are refused
 * (SCHIC_ERR_UNSUPPORTED: use schic_mh_fit_csr). */
int schic_mh_fit_csr_host(schic_mh_handle* h, int64_t n_rows, const void* indptr, int32_t indptr_bits, const void* indices,
                          int32_t index_bits, const void* data, int32_t data_kind, int32_t append, int32_t async_upload);

/* MinHash.fit on a CSR shard already in HBM. The arrays are BORROWED (not copied) and must stay
 * valid until the next fit or destroy. d_indptr [n_rows+1] int64 (absolute positions into
 * d_indices/d_data), d_indices uint32, d_data float32 (may be NULL when fast=1). append must be 0. */
int schic_mh_fit_csr_device(schic_mh_handle* h, int64_t n_rows, int64_t nnz, const int64_t* d_indptr,
                            const uint32_t* d_indices, const float* d_data, int32_t append);

/* The same with the value of d_indptr[0] given by the caller (0 for a stand-alone shard): nothing is read back, so
 * the call enqueues without draining the stream -- with schic_mh_set_feature_range a whole build then runs
 * without a single host synchronisation. */
int schic_mh_fit_csr_device_at(schic_mh_handle* h, int64_t n_rows, int64_t nnz, int64_t indptr0, const int64_t* d_indptr,
                               const uint32_t* d_indices, const float* d_data, int32_t append);

/* MinHash.kneighbors with results left in HBM: d_idx [n_local,k] int32, d_dist [n_local,k] float32,
 * d_nvalid [n_local] int32. Asynchronous on the handle's stream. */
int schic_mh_kneighbors_device(schic_mh_handle* h, int32_t k, int32_t fast, int32_t* d_idx, float* d_dist,
                               int32_t* d_nvalid);

/* MinHash.kneighbors_graph(mode='distance') (scHicClusterMinHash.py:355) with the CSR left in HBM: d_indptr
 * [n_local+1] int64, d_indices / d_data with room for n_local*k entries (the first d_indptr[n_local] are written).
 * Asynchronous on the handle's stream. */
int schic_mh_kneighbors_graph_device(schic_mh_handle* h, int32_t k, int32_t fast, int64_t* d_indptr, int32_t* d_indices,
                                     float* d_data);

/* scHicCluster -drm knn on device outputs: as schic_mh_kneighbors_exact (include/schic_mh.h), results left in HBM. */
int schic_mh_kneighbors_exact_device(schic_mh_handle* h, int32_t k, int32_t include_self, int32_t* d_idx, float* d_dist,
                                     int32_t* d_nvalid);

/* Device address of the local signature matrix (uint32 [n_local, H]) -- the send buffer of the
 * signature all-gather. */
int schic_mh_signatures_device(schic_mh_handle* h, void** d_sig_out, int64_t* n_rows_out);

/* Device addresses of the local CSR shard as the handle holds it (after a host or device fit):
 * the send buffers of the CSR all-gather that the distributed exact rerank needs. */
int schic_mh_csr_device(schic_mh_handle* h, const int64_t** d_indptr_out, const uint32_t** d_indices_out,
                        const float** d_data_out, int64_t* nnz_out);

/* Distributed query: use an external (all-gathered, borrowed) database of n_total rows; this
 * handle's local rows are database rows [row0, row0+n_local). d_sig_all uint32 [n_total, H];
 * the CSR of the whole database (d_indptr_all/d_indices_all/d_data_all) is needed only for the
 * exact rerank. Passing d_sig_all == NULL returns to the local database. */
int schic_mh_set_database_device(schic_mh_handle* h, int64_t n_total, int64_t row0, const uint32_t* d_sig_all,
                                 const int64_t* d_indptr_all, const uint32_t* d_indices_all,
                                 const float* d_data_all);

/* The CSR part of the database handed to schic_mh_set_database_device may still be in flight (e.g. an all-gather on
 * another stream): `cuda_event` (a cudaEvent_t recorded behind that transfer; NULL = none) is waited for on the
 * handle's stream right before the exact rerank touches the CSR arrays, so collision counting and the top-k select
 * overlap the transfer. The signature part must be complete in stream order as usual. Cleared by the next
 * schic_mh_set_database_device / fit. */
int schic_mh_set_database_ready_event(schic_mh_handle* h, void* cuda_event);

/* Per-row preparation of the exact rerank, normally computed by the library for the whole database on first use, in one
 * pass over ids + values: fp64 row norms, the window directory (offsets of 2^16-id windows in every row), the PACKED
 * STREAM of the counts kernel -- one uint32 per non-zero, (id mod 2^16) << 16 | count, same indexing as d_indices -- and
 * four flags {a value is not an integer in [0, 65535]; number of values >= 255; an id beyond the declared range; -}.
 * In a sharded run every rank prepares its OWN rows and the arrays are exchanged instead of being recomputed by everyone:
 *   schic_mh_rerank_windows      number of id windows for the range declared with schic_mh_set_feature_range
 *                                (0: no hint, or the range needs the hash-table kernel -- then do not use the two below)
 *   schic_mh_rerank_prep_rows    norms2[n_rows] (float64), dir[n_rows, n_windows+1] (uint32), packed[nnz] (uint32, may be
 *                                NULL), flags[4] (uint32, zeroed here) of the given CSR rows, queued on the handle's stream
 *   schic_mh_set_database_prep_device   norms / directory (and optionally packed stream + flags, OR-ed / summed over the
 *                                ranks, + total non-zeros) of ALL database rows for the next kneighbors call on an external
 *                                database (call after schic_mh_set_database_device, which clears them). The packed
 *                                stream must be 16-byte aligned and followed by >= 16 readable bytes. With a packed
 *                                stream the exact rerank does not need d_indices_all / d_data_all of
 *                                schic_mh_set_database_device (pass NULL: half the exchange volume) -- valid when the
 *                                values are small integer counts, which the flags verify: if they are not, the next
 *                                synchronising call reports SCHIC_ERR_INVALID_ARG instead of returning wrong distances.
 *   schic_mh_rerank_kernel_name  which kernel did the work of the last exact rerank ("k4c_rerank_kernel" = counts
 *                                kernel, "k4_rerank_window_kernel" / "k4_rerank_kernel" = generic float kernels);
 *                                out must hold 32 bytes. Synchronises the stream. */
int schic_mh_rerank_windows(schic_mh_handle* h, int32_t* n_windows_out);
int schic_mh_rerank_prep_rows(schic_mh_handle* h, int64_t n_rows, const int64_t* d_indptr, const uint32_t* d_indices,
                              const float* d_data, int32_t n_windows, double* d_norm2_out, uint32_t* d_dir_out,
                              uint32_t* d_packed_out, uint32_t* d_flags_out);
int schic_mh_set_database_prep_device(schic_mh_handle* h, const double* d_norm2_all, const uint32_t* d_dir_all,
                                      int32_t n_windows, const uint32_t* d_packed_all, const uint32_t* d_flags_all,
                                      int64_t nnz_total);
int schic_mh_rerank_kernel_name(schic_mh_handle* h, char* out, int32_t n_out);

/* Collision counts of a BLOCK of the (database x database) count matrix, for callers that assemble the matrix
 * themselves: rows [q_row0, q_row0 + q_rows) against columns [d_row0, d_row0 + d_rows) of the database (the external one
 * when set, else the fitted rows), written to d_cnt[(q - q_row0) * ld + (d - d_row0)] (uint16) and -- when d_cnt_t is
 * given -- transposed to d_cnt_t[(d - d_row0) * ld_t + (q - q_row0)]. Identical row and column ranges take the symmetric
 * kernel (half the compares). Rule 5 (max_bin_size) is applied as in kneighbors. Needs the packed K3 (H <= 4094, <= 57 344
 * database rows), else SCHIC_ERR_UNSUPPORTED. Asynchronous on the handle's stream.
 * Use: the count matrix is symmetric, so in a sharded run each (rank x rank) block needs computing only once: every
 * rank computes its diagonal block plus half of every off-diagonal block of its rows and sends the transposes to the
 * peers that own those columns' rows -- n^2 H / (2 G) compares per rank instead of n^2 H / G.
 * schic_mh_set_counts_device: the next kneighbors* call on the handle takes d_counts ([local rows, ld], ld a multiple of 8,
 * 16-byte aligned) as its collision counts instead of computing them (one-shot; NULL clears). */
int schic_mh_collision_block_device(schic_mh_handle* h, int64_t q_row0, int64_t q_rows, int64_t d_row0, int64_t d_rows,
                                    uint16_t* d_cnt, int64_t ld, uint16_t* d_cnt_t, int64_t ld_t);
int schic_mh_set_counts_device(schic_mh_handle* h, const uint16_t* d_counts, int64_t ld);

/* Run on the caller's stream (a cudaStream_t, e.g. torch's current stream) instead of the
 * handle's own. NULL selects the legacy default stream. */
int schic_mh_set_stream(schic_mh_handle* h, void* cuda_stream);
int schic_mh_synchronize(schic_mh_handle* h);

/* schic_mh_fit_csr with the values given as unsigned integer counts (count_bits = 8 or 16) instead of float32:
 * Hi-C values are contact counts, and narrow counts halve or quarter the PCIe bytes of the value array; they are
 * widened to float32 on the device, everything downstream is unchanged. async_upload as schic_mh_fit_csr_async. */
int schic_mh_fit_csr_counts(schic_mh_handle* h, int64_t n_rows, const int64_t* indptr, const void* indices,
                            int32_t index_bits, const void* counts, int32_t count_bits, int32_t append,
                            int32_t async_upload);

/* Optional hint: every feature id handed to fit is < n_features (the column count of the CSR matrix,
 * nbins^2 in the reference, utilities.py:131). With it the library sizes its per-id tables and the rerank's window
 * directory without reading the largest id back from the device, i.e. fit/kneighbors enqueue without a host
 * synchronisation. An id >= n_features is a caller error: it is detected on the device, never written out of
 * bounds, and reported by the next synchronising call (SCHIC_ERR_INVALID_ARG). 0 clears the hint. */
int schic_mh_set_feature_range(schic_mh_handle* h, uint64_t n_features);

/* CUDA-event time (ms) of the stages of the most recent fit + kneighbors on this handle:
 * [0] h2d, [1] signatures (K1), [2] collisions (K3), [3] select, [4] rerank (K4 launch alone), [5] graph (K5),
 * [6] d2h, [7] rerank preparation (row norms, window directory), [8] sum; n_out >= 9. Synchronises the stream. */
int schic_mh_stage_ms(schic_mh_handle* h, float* out, int32_t n_out);

/* How K1 of the most recent fit got its shared hash table (k1_table.cu): 0 no table (per-cell kernels), 1 built from the
 * batch's own ids, 2 built for every id of the range declared with schic_mh_set_feature_range, 3 the whole-range table
 * of an earlier fit on this handle re-used. Such a table depends only on (range, number of hash functions, hash width)
 * -- not on the data -- so the handle keeps it across fits of the same shape (repeated builds, partial_fit chunks,
 * every rank of a sharded run): the second fit of a shape builds it, later ones only look up. SCHIC_K1_NO_CACHE=1
 * (environment) disables the re-use. Results are bit-identical either way. */
int schic_mh_k1_table_state(schic_mh_handle* h, int32_t* state_out);

/* Number of kernels this handle has launched so far. */
int schic_mh_launch_count(schic_mh_handle* h, int64_t* n_out);

/* INT32 issue-rate micro-benchmark, Tint-op/s: out[0] LOP3, [1] IMAD, [2] 1:1 mix, [3] SM MHz,
 * [4] SHF, [5] IMAD.HI, [6] VIMNMX, [7] VIMNMX3, [8] IADD, [9] IMAD.WIDE, [10] HSET2, [11] HADD2,
 * [12] 1:1 HSET2+HADD2, [13] ISETP + predicated FADD pairs; out must hold 16 doubles. */
int schic_int32_peak(int32_t device, int32_t iters, double* out);

/* Experiment hook: K1 with instruction-selection variant v on device-resident CSR; elapsed ms. */
int schic_k1_variant_ms(int32_t variant, int64_t n_rows, int64_t nnz, const int64_t* d_indptr,
                        const uint32_t* d_indices, int32_t n_hash, uint32_t* d_sig, int32_t reps, float* ms_out);

/* Page-locked host memory for the e2e path (numpy arrays are built over it). */
int schic_pinned_alloc(int64_t bytes, void** out);
int schic_pinned_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* SCHIC_MH_CUDA_H_ */
