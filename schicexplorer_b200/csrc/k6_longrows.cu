// k6_longrows.cu -- rows too long for the shared-memory sorts of select / graph emit / rerank merge, sm_100a.
//
// The reference tool's default is a DENSE kNN graph: scHicClusterMinHash sets k = number of cells when -k is not given
// (/root/reference/schicexplorer/scHicClusterMinHash.py:255-257), so every per-row stage sees rows of n entries. The
// regular kernels order a row inside one CTA's shared memory (k3_select.cu, k5_graph.cu, k4_merge_rows_kernel), which
// caps a row at 16 384 - 32 768 keys. Beyond that these kernels take over: the same packed 64-bit keys, ordered by a
// bitonic sort that works in GLOBAL memory, one CTA per row (a 65 536-key row is 512 KB: the 148 rows in flight stay in
// L2). Slower per key than the shared-memory path, but a 50 000-cell dense graph needs two such sorts per row and
// still finishes in a fraction of a second; results are identical by construction (same keys, same order).
#include "common.cuh"
#include "sort.cuh"

namespace schic {

constexpr int LR_THREADS = 1024;

// in-place ascending bitonic sort of each row's P (power of two) keys in global memory
__global__ void __launch_bounds__(LR_THREADS)
longrow_sort_kernel(unsigned long long* __restrict__ keys, int P) {
    unsigned long long* k = keys + (int64_t)blockIdx.x * P;
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += LR_THREADS) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = k[lo], b = k[hi];
                if ((a > b) == up) { k[lo] = b; k[hi] = a; }
            }
        }
    }
}

// ---- select: (count desc, index asc) keys of one row of the count matrix ---------------------------------------------
__global__ void __launch_bounds__(256)
longrow_select_keys_kernel(const uint16_t* __restrict__ cnt, int64_t ld, int64_t ndb, int H, int min_blocks,
                           unsigned long long* __restrict__ keys, int P) {
    const int64_t q = blockIdx.x;
    const uint16_t* __restrict__ row = cnt + q * ld;
    const int thr = max(min_blocks, 1);
    for (int d = threadIdx.x; d < P; d += blockDim.x) {
        unsigned long long key = ~0ull;
        if (d < ndb) {
            const int c = row[d];
            if (c >= thr) key = ((unsigned long long)(uint32_t)(H - c) << 32) | (unsigned long long)(uint32_t)d;
        }
        keys[q * (int64_t)P + d] = key;
    }
}

__global__ void __launch_bounds__(256)
longrow_select_take_kernel(const unsigned long long* __restrict__ keys, int P, int kcand, int H, int32_t* __restrict__ idx,
                           int32_t* __restrict__ cntsel, int32_t* __restrict__ nvalid) {
    const int64_t q = blockIdx.x;
    const unsigned long long* __restrict__ k = keys + q * (int64_t)P;
    int nv_local = 0;
    for (int t = threadIdx.x; t < kcand; t += blockDim.x) {
        const unsigned long long key = t < P ? k[t] : ~0ull;
        const bool ok = key != ~0ull;
        idx[q * (int64_t)kcand + t] = ok ? (int32_t)(uint32_t)key : -1;
        cntsel[q * (int64_t)kcand + t] = ok ? H - (int32_t)(uint32_t)(key >> 32) : 0;
        nv_local += ok ? 1 : 0;
    }
    __shared__ int s_nv;
    if (threadIdx.x == 0) s_nv = 0;
    __syncthreads();
    if (nv_local) atomicAdd(&s_nv, nv_local);
    __syncthreads();
    if (threadIdx.x == 0) nvalid[q] = s_nv;
}

int longrow_pow2(int64_t n) { int p = 1; while (p < n) p <<= 1; return p; }

// Same contract as launch_select_topk for rows [0, nq) of d_cnt. d_keys: nq_block * P uint64 of scratch, P =
// longrow_pow2(ndb); the rows are processed nq_block at a time.
int launch_select_topk_long(const uint16_t* d_cnt, int64_t nq, int64_t ndb, int64_t ld_cnt, int H, int min_blocks, int kcand,
                            int32_t* d_idx, int32_t* d_cntsel, int32_t* d_nvalid, unsigned long long* d_keys,
                            int64_t nq_block, cudaStream_t stream) {
    if (nq <= 0) return 0;
    const int P = longrow_pow2(ndb);
    int nl = 0;
    for (int64_t q0 = 0; q0 < nq; q0 += nq_block) {
        const unsigned nb = (unsigned)std::min<int64_t>(nq_block, nq - q0);
        longrow_select_keys_kernel<<<nb, 256, 0, stream>>>(d_cnt + q0 * ld_cnt, ld_cnt, ndb, H, min_blocks, d_keys, P);
        longrow_sort_kernel<<<nb, LR_THREADS, 0, stream>>>(d_keys, P);
        longrow_select_take_kernel<<<nb, 256, 0, stream>>>(d_keys, P, kcand, H, d_idx + q0 * kcand, d_cntsel + q0 * kcand,
                                                           d_nvalid + q0);
        nl += 3;
    }
    return nl;
}

// ---- graph emit: rows ordered by column -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
longrow_emit_keys_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dist, const int32_t* __restrict__ nvalid,
                         int k, unsigned long long* __restrict__ keys, int P) {
    const int64_t q = blockIdx.x;
    const int nv = nvalid[q];
    for (int t = threadIdx.x; t < P; t += blockDim.x) {
        unsigned long long key = ~0ull;
        if (t < nv) key = ((unsigned long long)(uint32_t)idx[q * (int64_t)k + t] << 32) | __float_as_uint(dist[q * (int64_t)k + t]);
        keys[q * (int64_t)P + t] = key;
    }
}

__global__ void __launch_bounds__(256)
longrow_emit_take_kernel(const unsigned long long* __restrict__ keys, int P, const int32_t* __restrict__ nvalid,
                         const int64_t* __restrict__ indptr, int32_t* __restrict__ indices, float* __restrict__ data) {
    const int64_t q = blockIdx.x;
    const int nv = nvalid[q];
    const int64_t o = indptr[q];
    for (int t = threadIdx.x; t < nv; t += blockDim.x) {
        const unsigned long long key = keys[q * (int64_t)P + t];
        indices[o + t] = (int32_t)(uint32_t)(key >> 32);
        data[o + t] = __uint_as_float((uint32_t)(key & 0xFFFFFFFFull));
    }
}

// Same contract as launch_graph_emit.
int launch_graph_emit_long(const int32_t* d_idx, const float* d_dist, const int32_t* d_nvalid, int64_t nq, int k,
                           int64_t* d_indptr, int32_t* d_indices, float* d_data, unsigned long long* d_keys,
                           int64_t nq_block, cudaStream_t stream) {
    if (nq <= 0) return 0;
    const int P = longrow_pow2(k);
    int nl = launch_scan_nvalid(d_nvalid, nq, d_indptr, stream);
    for (int64_t q0 = 0; q0 < nq; q0 += nq_block) {
        const unsigned nb = (unsigned)std::min<int64_t>(nq_block, nq - q0);
        longrow_emit_keys_kernel<<<nb, 256, 0, stream>>>(d_idx + q0 * k, d_dist + q0 * k, d_nvalid + q0, k, d_keys, P);
        longrow_sort_kernel<<<nb, LR_THREADS, 0, stream>>>(d_keys, P);
        longrow_emit_take_kernel<<<nb, 256, 0, stream>>>(d_keys, P, d_nvalid + q0, d_indptr + q0, d_indices, d_data);
        nl += 3;
    }
    return nl;
}

// ---- rerank merge: the chunk-sorted (distance, index) pairs of a row ---------------------------------------------------------
__global__ void __launch_bounds__(256)
longrow_merge_keys_kernel(const int32_t* __restrict__ part_idx, const float* __restrict__ part_dist, int part_stride,
                          int64_t exclude_row0, unsigned long long* __restrict__ keys, int P) {
    const int64_t q = blockIdx.x;
    const int64_t self = exclude_row0 >= 0 ? exclude_row0 + q : -1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < part_stride) {
            const int32_t d = part_idx[q * (int64_t)part_stride + i];
            if (d >= 0 && d != self)
                key = ((unsigned long long)__float_as_uint(part_dist[q * (int64_t)part_stride + i]) << 32) | (uint32_t)d;
        }
        keys[q * (int64_t)P + i] = key;
    }
}

__global__ void __launch_bounds__(256)
longrow_merge_take_kernel(const unsigned long long* __restrict__ keys, int P, const int32_t* __restrict__ cand_nvalid, int kcand,
                          int n_all, int64_t exclude_row0, int k, int32_t* __restrict__ idx_out, float* __restrict__ dist_out,
                          int32_t* __restrict__ nvalid_out) {
    const int64_t q = blockIdx.x;
    const int64_t self = exclude_row0 >= 0 ? exclude_row0 + q : -1;
    const int total = cand_nvalid ? min(cand_nvalid[q], kcand) : n_all - (self >= 0 && self < n_all ? 1 : 0);
    const int nv = min(total, k);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const unsigned long long key = keys[q * (int64_t)P + (t < P ? t : P - 1)];
        const bool ok = t < nv && t < P && key != ~0ull;
        idx_out[q * (int64_t)k + t] = ok ? (int32_t)(uint32_t)key : -1;
        dist_out[q * (int64_t)k + t] = ok ? __uint_as_float((uint32_t)(key >> 32)) : INFINITY;
    }
    if (threadIdx.x == 0) nvalid_out[q] = nv;
}

// Same contract as launch_rerank_merge (outputs of row q at d_*_out[q * k ...]).
int launch_rerank_merge_long(const int32_t* d_part_idx, const float* d_part_dist, int part_stride, const int32_t* d_cand_nvalid,
                             int kcand, int n_all, int64_t exclude_row0, int64_t nq, int k, int32_t* d_idx_out,
                             float* d_dist_out, int32_t* d_nvalid_out, unsigned long long* d_keys, int64_t nq_block,
                             cudaStream_t stream) {
    if (nq <= 0) return 0;
    const int P = longrow_pow2(part_stride);
    int nl = 0;
    for (int64_t q0 = 0; q0 < nq; q0 += nq_block) {
        const unsigned nb = (unsigned)std::min<int64_t>(nq_block, nq - q0);
        longrow_merge_keys_kernel<<<nb, 256, 0, stream>>>(d_part_idx + q0 * part_stride, d_part_dist + q0 * part_stride,
                                                          part_stride, exclude_row0 >= 0 ? exclude_row0 + q0 : -1, d_keys, P);
        longrow_sort_kernel<<<nb, LR_THREADS, 0, stream>>>(d_keys, P);
        longrow_merge_take_kernel<<<nb, 256, 0, stream>>>(d_keys, P, d_cand_nvalid ? d_cand_nvalid + q0 : nullptr, kcand, n_all,
                                                          exclude_row0 >= 0 ? exclude_row0 + q0 : -1, k, d_idx_out + q0 * k,
                                                          d_dist_out + q0 * k, d_nvalid_out + q0);
        nl += 3;
    }
    return nl;
}

}  // namespace schic
