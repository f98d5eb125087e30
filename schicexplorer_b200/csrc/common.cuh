// common.cuh -- shared declarations of the sm_100a MinHash kNN kernels (host-side launchers).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/schic_mh.h"

namespace schic {

// Sets the thread-local message schic_mh_last_error() returns (capi.cu); returns code.
int set_error(int code, const std::string& msg);

constexpr uint32_t kModulus = SCHIC_MH_MODULUS;  // 2^31 - 1; also the empty-row sentinel

// ---- K1: MinHash signatures (k1_signatures.cu) ------------------------------------------
// sig[row_offset + r][j] = min over the row's features f of wang32((f+1)(j+1)) mod M.
// indptr [n_rows+1] (device, indptr[0] may be non-zero), feat = feature ids (device).
// Launches a fill (sentinel) kernel and the hashing kernel on `stream`; returns launch count.
int launch_signatures32(const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows, int64_t nnz,
                        int64_t indptr0, int H, uint32_t* d_sig /* first row of this batch */,
                        cudaStream_t stream);
// Filtered variant (whole rows per CTA, cheap "cannot be a new minimum" test per evaluation, exact path
// only when it fails). d_scratch: k1f_scratch_bytes(n_rows, nnz) bytes of device memory.
size_t k1f_scratch_bytes(int64_t n_rows, int64_t nnz);
int launch_signatures32_filter(const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows, int64_t nnz,
                               int H, uint32_t* d_sig, void* d_scratch, int direct /* 1: no shared-memory tile */,
                               cudaStream_t stream);
// Shared-table variant (k1_table.cu): every (feature id, hash function) is evaluated once for the whole batch.
bool k1t_applicable(int64_t n_rows, int64_t nnz, uint32_t max_id, int H);
double k1t_tau(int64_t n_rows, int64_t nnz, int H);
size_t k1t_entries_capacity(int64_t n_rows, int64_t nnz, uint32_t max_id, int H);
size_t k1t_items_bytes(int64_t n_rows, int64_t nnz);
int launch_max_id(const uint32_t* d_feat, int64_t nnz, uint32_t* d_out_max, cudaStream_t stream);
int launch_signatures_table(int width /* 32 | 64: HashSpec */, const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows, int64_t nnz,
                              int64_t indptr0, uint32_t max_id, int H, uint32_t* d_sig, uint8_t* d_present,
                              uint2* d_slot, uint16_t* d_entries, size_t entries_cap, uint32_t* d_counters,
                              void* d_items, int phase /* 0 all | 1 build for every id | 2 lookup */, cudaStream_t stream);
// variant >= 0 selects an instruction-selection variant (only in builds with -DK1_ALL_VARIANTS).
int launch_signatures32_variant(int variant, const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows,
                                int64_t nnz, int64_t indptr0, int H, uint32_t* d_sig, cudaStream_t stream);
int launch_signatures64(const int64_t* d_indptr, const uint32_t* d_feat_lo, const uint32_t* d_feat_hi,
                        int64_t n_rows, int64_t nnz, int64_t indptr0, int H, uint32_t* d_sig,
                        cudaStream_t stream);

// ---- K2: optional block combine ("shingle", k2_combine.cu) ------------------------------------
// keys[r][b] = fold of raw[r][b*block .. b*block+block-1] (rule 4); raw [n, H_raw] -> keys [n, H_raw / block].
int launch_block_combine(int width, const uint32_t* d_raw, int64_t n, int H_raw, int block, uint32_t* d_keys, cudaStream_t stream);

// ---- K2b: max_bin_size bucket pruning (k2_buckets.cu) -----------------------------------------
// out[c][j] = sig[c][j] unless its bucket {d : sig[d][j] == sig[c][j]} has >= max_bin_size members, then 0.
int launch_bucket_mask(const uint32_t* d_sig, int64_t n, int H, int64_t max_bin_size, uint32_t* d_out,
                       cudaStream_t stream);

// ---- K3: collision counts (k3_collisions.cu) ---------------------------------------------
// cnt[q][d] = #{j : sigQ[q][j] == sigD[d][j], value admissible}; uint16 row major [nq, ld_out].
int launch_collision_counts(const uint32_t* d_sig_q, int64_t nq, const uint32_t* d_sig_db, int64_t ndb,
                            int H, uint16_t* d_cnt, int64_t ld_out, cudaStream_t stream);

// Packed variant: signatures renamed per hash function to 16-bit ids stored as normal fp16 bit patterns; one
// HSET2 + HADD2 pair handles two hash functions. Same counts, bit for bit. H <= 4094, ndb <= 57 344.
bool k3h_applicable(int64_t ndb, int H);
int k3h_padded_hashes(int H);
size_t k3h_table_elems(int64_t ndb, int H);
int launch_rename_signatures(const uint32_t* d_sig_db, int64_t ndb, int H, uint32_t* d_table, uint16_t* d_sig16,
                             uint32_t* d_sizes /* nullable */, int64_t max_bin_size /* 0: no pruning */, cudaStream_t stream);
int launch_collision_counts_packed(const uint16_t* d_q16, int64_t nq, const uint16_t* d_d16, int64_t ndb, int H,
                                   uint16_t* d_cnt, int64_t ld_out, cudaStream_t stream);
// ... plus the transposed block: d_cnt_t [ndb, ld_t], cnt_t[d][q] = cnt[q][d] (what the peer rank that owns rows d needs)
int launch_collision_counts_packed_t(const uint16_t* d_q16, int64_t nq, const uint16_t* d_d16, int64_t ndb, int H,
                                     uint16_t* d_cnt, int64_t ld_out, uint16_t* d_cnt_t, int64_t ld_t, cudaStream_t stream);

// TMA-staged variant: the renamed signatures transposed (sig16T, k3t_transposed_elems uint32), a tensor map over them
// (128 bytes, 64-byte aligned host memory), counts of database rows [q_row0, +nq) x [d_row0, +ndb).
size_t k3t_transposed_elems(int64_t n, int H);
int launch_transpose_signatures(const uint16_t* d_sig16, int64_t n, int H, uint32_t* d_sig16T, cudaStream_t stream);
int k3t_make_tensor_map(const uint32_t* d_sig16T, int64_t n, int H, void* out_map);
int launch_collision_counts_tma(const void* map, int64_t q_row0, int64_t nq, int64_t d_row0, int64_t ndb, int H, uint16_t* d_cnt,
                                int64_t ld_out, uint16_t* d_cnt_t, int64_t ld_t, cudaStream_t stream);

// ---- K3b: threshold + top-k select + fast distance (k3_select.cu) --------------------------
// For each query row: candidates with cnt >= min_blocks (and > 0), ordered by (cnt desc, idx asc),
// first kcand kept. Writes idx [nq,kcand] (-1 padded), cnt [nq,kcand], n_valid [nq].
// d_row_flags: device int [nq] scratch or nullptr. Non-null lets the launcher use the single-histogram kernel
// (H <= 4095); rows it cannot finish (more ties at the threshold than its buffer holds) are flagged there and
// redone by the two-level radix kernel in the same call.
int launch_select_topk(const uint16_t* d_cnt, int64_t nq, int64_t ndb, int64_t ld_cnt, int H,
                       int min_blocks, int kcand, int32_t* d_idx, int32_t* d_cntsel, int32_t* d_nvalid,
                       int* d_row_flags, cudaStream_t stream);
// dist = absolute ? cnt : 1 - cnt/H (fp32), first k of kcand columns; pads +inf / -1.
int launch_fast_distance(const int32_t* d_idx_in, const int32_t* d_cnt_in, const int32_t* d_nvalid_in,
                         int64_t nq, int kcand, int k, int H, int absolute, int32_t* d_idx_out,
                         float* d_dist_out, int32_t* d_nvalid_out, cudaStream_t stream);

// ---- K4: exact euclidean rerank (k4_rerank.cu, k4c_rerank.cu) ---------------------------------
// id windows of the rerank: the per-row window directory is built for windows of 2^SCHIC_K4C_WIN_BITS ids (the byte
// table of the counts kernel); the generic windowed kernel works on 2^19-id windows = 4 of them.
#ifndef SCHIC_K4C_WIN_BITS
#define SCHIC_K4C_WIN_BITS 16
#endif
// Per-fit preparation in one pass over ids + values: fp64 row norms, window directory dir[row][0..n_windows] (uint32
// offsets from the row start), the packed stream ((id mod 2^16) << 16 | count, indexed like d_feat; nullable) and
// d_flags[4] (accumulated, zero them first): [0] a value is not an integer in [0, 65535], [1] number of values >= 255,
// [2] an id lies beyond the last window.
int k4c_window_count(uint32_t max_feature);
int launch_rerank_prep(const int64_t* d_indptr, const uint32_t* d_feat, const float* d_val, int64_t n_rows, int n_windows,
                       double* d_norm2, uint32_t* d_dir, uint32_t* d_packed, uint32_t* d_flags, cudaStream_t stream);
// Counts kernel (packed stream + byte table + window-group-major grid) and its finalize; output contract of
// launch_rerank. Both exit at once unless d_flags says the data is packable (flags[0] == 0 && flags[1] <= esc_limit).
// d_dotacc: k4c_dotacc_elems(nq, kcand) uint64 of scratch. group_bytes: share of the packed stream one window group
// may span (sized for L2; <= 0: one group). Returns the number of launches, -1 when kcand does not fit.
size_t k4c_dotacc_elems(int64_t nq, int kcand);
bool k4c_direct_selected();
const char* k4c_kernel_name();   // "k4c_rerank_direct_kernel" | "k4c_rerank_kernel" (SCHIC_K4_KERNEL=iter)
int launch_rerank_counts(const int64_t* d_indptr, const uint32_t* d_packed, const double* d_norm2, const uint32_t* d_dir,
                         int n_windows, int64_t db_nnz, int64_t group_bytes, int64_t q_row0, int64_t nq,
                         const int32_t* d_cand_idx, const int32_t* d_cand_nvalid, int cand_stride, int cand_off, int kcand,
                         int k, int32_t* d_idx_out, float* d_dist_out, int out_stride, int out_off, int final,
                         int32_t* d_nvalid_out, int n_all, const uint32_t* d_flags, uint32_t esc_limit,
                         unsigned long long* d_dotacc, cudaStream_t stream);
// uint16 / uint8 counts -> float32 values (host uploads of integer counts cross PCIe narrow)
int launch_widen_counts(const void* d_src, int bits, int64_t n, float* d_dst, cudaStream_t stream);
// Window directory of the windowed rerank kernel: largest feature id of the database (device
// scalar), number of windows for it, and dir[row][0..n_windows] (uint32 offsets from the row start).
int launch_row_max_feature(const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows, uint32_t* d_out_max,
                           cudaStream_t stream);
// For query q (database row q_row0+q) and each of its candidates: euclidean distance over the
// union of supports; then order by (dist asc, idx asc), keep k. CSR arrays are the database's.
// d_dir != nullptr selects the windowed kernel (needs the directory of launch_rerank_prep: n_windows counts its
// 2^SCHIC_K4C_WIN_BITS-id windows), else the hash-table kernel. d_skip_flags != nullptr: the kernels exit at once when
// the flags say the counts kernel (launched before them) has done the work.
// A launch handles candidate columns [cand_off, cand_off+kcand) of lists with row stride cand_stride and writes the k
// best (sorted) to output columns [out_off, out_off+k) of arrays with row stride out_stride; final != 0 also writes
// d_nvalid_out. Returns -1 when kcand does not fit one CTA's shared memory (the caller then chunks + merges).
int launch_rerank(const int64_t* d_indptr, const uint32_t* d_feat, const float* d_val, int64_t indptr0,
                  const double* d_norm2, const uint32_t* d_dir, int n_windows, int window_mode /* 32 | 64: entry width */,
                  int64_t q_row0, int64_t nq, const int32_t* d_cand_idx, const int32_t* d_cand_nvalid, int cand_stride,
                  int cand_off, int kcand, int k, int32_t* d_idx_out, float* d_dist_out, int out_stride, int out_off,
                  int final, int32_t* d_nvalid_out, int n_all /* d_cand_idx == nullptr: candidates = rows [0, n_all) */,
                  int allow_hash /* 0: windowed kernel or -1 */, const uint32_t* d_skip_flags, uint32_t esc_limit,
                  cudaStream_t stream);
// Per row: sort the [part_stride] chunk-sorted (distance, index) pairs by (dist asc, idx asc), keep k, write n_valid
// (= min(k, list length); list length = min(cand_nvalid, kcand), or n_all [- 1 if exclude_row0 >= 0] for brute force).
int launch_rerank_merge(const int32_t* d_part_idx, const float* d_part_dist, int part_stride,
                        const int32_t* d_cand_nvalid, int kcand, int n_all, int64_t exclude_row0, int64_t nq, int k,
                        int32_t* d_idx_out, float* d_dist_out, int32_t* d_nvalid_out, cudaStream_t stream);

// ---- K5: graph emit (k5_graph.cu) ---------------------------------------------------------
// CSR of the kNN graph: indptr = exclusive scan of n_valid, rows sorted by column.
int launch_graph_emit(const int32_t* d_idx, const float* d_dist, const int32_t* d_nvalid, int64_t nq, int k,
                      int64_t* d_indptr, int32_t* d_indices, float* d_data, cudaStream_t stream);

int launch_scan_nvalid(const int32_t* d_nvalid, int64_t nq, int64_t* d_indptr, cudaStream_t stream);

// ---- rows beyond the shared-memory sorts (k6_longrows.cu): same contracts as launch_select_topk / launch_graph_emit /
// launch_rerank_merge, ordered by a global-memory bitonic sort. d_keys: nq_block * longrow_pow2(row length) uint64.
int longrow_pow2(int64_t n);
int launch_select_topk_long(const uint16_t* d_cnt, int64_t nq, int64_t ndb, int64_t ld_cnt, int H, int min_blocks, int kcand,
                            int32_t* d_idx, int32_t* d_cntsel, int32_t* d_nvalid, unsigned long long* d_keys,
                            int64_t nq_block, cudaStream_t stream);
int launch_graph_emit_long(const int32_t* d_idx, const float* d_dist, const int32_t* d_nvalid, int64_t nq, int k,
                           int64_t* d_indptr, int32_t* d_indices, float* d_data, unsigned long long* d_keys,
                           int64_t nq_block, cudaStream_t stream);
int launch_rerank_merge_long(const int32_t* d_part_idx, const float* d_part_dist, int part_stride, const int32_t* d_cand_nvalid,
                             int kcand, int n_all, int64_t exclude_row0, int64_t nq, int k, int32_t* d_idx_out,
                             float* d_dist_out, int32_t* d_nvalid_out, unsigned long long* d_keys, int64_t nq_block,
                             cudaStream_t stream);

// ---- ids beyond 32 bits for the rerank (k7_relabel.cu): labels[i] = rank of (hi[i] << 32 | lo[i]) among the distinct ids
size_t relabel_scratch_bytes(int64_t n);
int launch_relabel_ids(const uint32_t* d_lo, const uint32_t* d_hi, int64_t n, uint32_t* d_labels, uint32_t* d_max_label,
                       void* d_scratch, size_t scratch_bytes, cudaStream_t stream);

// ---- INT32 throughput micro-benchmark (int_peak.cu) -----------------------------------------
// out[0] ALU-pipe (LOP3/SHF/IADD3) Tint-op/s, out[1] FMA-pipe (IMAD), out[2] 1:1 mix, out[3] SM MHz.
int run_int32_peak(int iters, double* out4, cudaStream_t stream);

}  // namespace schic
