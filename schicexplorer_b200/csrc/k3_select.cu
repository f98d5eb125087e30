// k3_select.cu -- K3b: candidate filter, exact top-k by collision count, fast distance. sm_100a.
//
// Replaces the per-query "drop < minimal_blocks_in_common, sort by collisions, keep
// n_neighbors * excess_factor" step of sparse_neighbors_search.MinHash.kneighbors and its
// fast=True distance (kwargs at /root/reference/schicexplorer/scHicClusterMinHash.py:402-404;
// SURVEY 8(c) rules 7-8). Result order is (count desc, index asc) -- deterministic, where the
// reference's tie order depends on hash-map iteration.
//
// One CTA per query row of the uint16 count matrix. The k-th largest admissible count is found
// exactly with a two-level (high byte / low byte) radix histogram in shared memory -- three
// coalesced passes over the row, independent of H -- then the winners are compacted in index
// order (so ties at the threshold keep the lowest indices) and the <= kcand survivors are put in
// final order by a small in-shared-memory bitonic sort on packed (H - count, index) keys.
#include "common.cuh"
#include "sort.cuh"

namespace schic {

constexpr int SEL_THREADS = 256;

// keys: KeyT = uint32_t when index bits + count bits <= 32, else uint64_t
template <typename KeyT>
__global__ void __launch_bounds__(SEL_THREADS)
select_topk_kernel(const uint16_t* __restrict__ cnt, int64_t ndb, int64_t ld_cnt, int H, int min_blocks, int kcand,
                   int P /* next_pow2(kcand) */, int idx_bits, int32_t* __restrict__ idx_out,
                   int32_t* __restrict__ cnt_out, int32_t* __restrict__ nvalid_out,
                   const int* __restrict__ only_rows /* nullable: process row q only if only_rows[q] != 0 */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (only_rows && only_rows[blockIdx.x] == 0) return;
    KeyT* keys = reinterpret_cast<KeyT*>(smem_raw);
    __shared__ int hist[256];
    __shared__ int scan_scratch[33];
    __shared__ int s_bin, s_above, s_n_out;

    const int64_t q = blockIdx.x;
    const uint16_t* __restrict__ row = cnt + q * ld_cnt;
    const int lo_ok = min_blocks > 1 ? min_blocks : 1;   // admissible: cnt >= max(min_blocks, 1)

    // ---- pass 1: histogram of the high byte --------------------------------------------
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < ndb; i += blockDim.x) {
        const int c = row[i];
        if (c >= lo_ok) atomicAdd(&hist[c >> 8], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int above = 0, b = 255;
        for (; b >= 0; --b) {
            if (above + hist[b] >= kcand) break;
            above += hist[b];
        }
        s_bin = b;        // -1: fewer than kcand admissible entries -> take them all
        s_above = above;  // entries in higher bins
    }
    __syncthreads();
    const int hi_bin = s_bin;
    int thr;        // counts > thr are all taken; counts == thr taken lowest index first
    int need_eq;    // how many of the == thr entries to take
    if (hi_bin < 0) {
        thr = lo_ok - 1;
        need_eq = 0;
    } else {
        // ---- pass 2: histogram of the low byte inside the threshold bin -------------------
        const int above_hi = s_above;
        __syncthreads();
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int64_t i = threadIdx.x; i < ndb; i += blockDim.x) {
            const int c = row[i];
            if (c >= lo_ok && (c >> 8) == hi_bin) atomicAdd(&hist[c & 255], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int above = above_hi, b = 255;
            for (; b >= 0; --b) {
                if (above + hist[b] >= kcand) break;
                above += hist[b];
            }
            s_bin = b;
            s_above = above;
        }
        __syncthreads();
        thr = (hi_bin << 8) | s_bin;
        need_eq = kcand - s_above;
    }

    // ---- pass 3: compaction in index order ------------------------------------------------
    for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = ~(KeyT)0;
    if (threadIdx.x == 0) s_n_out = 0;
    __syncthreads();
    int eq_taken = 0;    // == thr entries accepted so far (uniform across the CTA)
    int n_out = 0;
    for (int64_t base = 0; base < ndb; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const int c = i < ndb ? (int)row[i] : 0;
        const bool gt = c > thr && c >= lo_ok;
        const bool eq = (c == thr) && need_eq > 0;
        int tot_eq;
        const int eq_rank = block_exclusive_scan(eq ? 1 : 0, scan_scratch, &tot_eq);
        const bool take = gt || (eq && (eq_taken + eq_rank) < need_eq);
        int tot_take;
        const int pos = block_exclusive_scan(take ? 1 : 0, scan_scratch, &tot_take);
        if (take) keys[n_out + pos] = ((KeyT)(uint32_t)(H - c) << idx_bits) | (KeyT)(uint32_t)i;
        eq_taken += tot_eq;
        n_out += tot_take;
    }
    __syncthreads();

    // ---- final order: (count desc, index asc) == ascending packed key ------------------------
    bitonic_sort_shared(keys, P);
    const KeyT idx_mask = (((KeyT)1) << idx_bits) - 1;
    for (int t = threadIdx.x; t < kcand; t += blockDim.x) {
        int32_t oi = -1, oc = 0;
        if (t < n_out) {
            const KeyT k = keys[t];
            oi = (int32_t)(k & idx_mask);
            oc = H - (int32_t)(k >> idx_bits);
        }
        idx_out[q * kcand + t] = oi;
        cnt_out[q * kcand + t] = oc;
    }
    if (threadIdx.x == 0) nvalid_out[q] = n_out;
}

// ---- single-histogram variant (H <= SELF_MAXH) ----------------------------------------------------
// The two-level radix kernel above reads the row three times and funnels every admissible entry into
// one or two histogram bins (counts below 256 share the high byte) -- at 50 k cells it took 13 ms.
// Counts are at most H, so for the hash counts in use a full histogram (one bin per count value)
// fits in shared memory: one pass finds the exact k-th largest count, a second pass appends the
// winners (counts above the threshold in any order -- the final sort orders them -- and the ties at
// the threshold through a small buffer that is sorted by index so the lowest indices win; a row with more
// ties than the buffer holds is flagged in overflow[] and redone by the radix kernel).
// Rows are read with 128-bit loads, 8 counts per thread.
constexpr int SELF_MAXH = 4095;
constexpr int SELF_EQCAP = 2048;

template <typename KeyT>
__global__ void __launch_bounds__(SEL_THREADS)
select_topk_fullhist_kernel(const uint16_t* __restrict__ cnt, int64_t ndb, int64_t ld_cnt, int H, int min_blocks,
                            int kcand, int P, int idx_bits, int32_t* __restrict__ idx_out,
                            int32_t* __restrict__ cnt_out, int32_t* __restrict__ nvalid_out, int* __restrict__ overflow /* [rows] */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KeyT* keys = reinterpret_cast<KeyT*>(smem_raw);                    // [P]
    int* hist = reinterpret_cast<int*>(keys + P);                      // [nbins], nbins = H + 1 rounded up to 256
    uint32_t* eqbuf = reinterpret_cast<uint32_t*>(hist + ((H + 256) & ~255));   // [SELF_EQCAP]
    __shared__ int scan_scratch[33];
    __shared__ int s_thr, s_need_eq, s_n_out, s_n_eq;

    const int64_t q = blockIdx.x;
    const uint16_t* __restrict__ row = cnt + q * ld_cnt;
    const int lo_ok = min_blocks > 1 ? min_blocks : 1;
    const int nbins = (H + 256) & ~255;
    const int64_t nvec = (ndb + 7) >> 3;

    for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = ~(KeyT)0;
    if (threadIdx.x == 0) { s_n_out = 0; s_n_eq = 0; }
    __syncthreads();
    // ---- pass 1: full histogram of the admissible counts -------------------------------------
    for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(row) + v);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = (int)((ws[u >> 1] >> ((u & 1) * 16)) & 0xFFFFu);
            if (c >= lo_ok && v * 8 + u < ndb) atomicAdd(&hist[c], 1);
        }
    }
    __syncthreads();
    // ---- threshold: largest t with #(count >= t) >= kcand (suffix sums over the bins) -------------
    {
        const int per = nbins / SEL_THREADS;                  // bins per thread, thread 0 owns the highest counts
        const int top = nbins - 1 - threadIdx.x * per;        // my bins: top, top-1, ..., top-per+1
        int mine = 0;
        for (int b = 0; b < per; ++b) mine += hist[top - b];
        int total;
        const int above = block_exclusive_scan(mine, scan_scratch, &total);   // entries in higher bins
        if (threadIdx.x == 0 && total < kcand) { s_thr = lo_ok - 1; s_need_eq = 0; }   // take every admissible entry
        if (total >= kcand && above < kcand && above + mine >= kcand) {       // the crossing lies in my bins
            int acc = above;
            for (int b = 0; b < per; ++b) {
                const int hb = hist[top - b];
                if (acc + hb >= kcand) { s_thr = top - b; s_need_eq = kcand - acc; break; }
                acc += hb;
            }
        }
    }
    __syncthreads();
    const int thr = s_thr, need_eq = s_need_eq;
    // ---- pass 2: append ------------------------------------------------------------------------
    for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(row) + v);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = (int)((ws[u >> 1] >> ((u & 1) * 16)) & 0xFFFFu);
            const int64_t i = v * 8 + u;
            if (i >= ndb || c < lo_ok) continue;
            if (c > thr) {
                const int pos = atomicAdd(&s_n_out, 1);
                keys[pos] = ((KeyT)(uint32_t)(H - c) << idx_bits) | (KeyT)(uint32_t)i;
            } else if (c == thr && need_eq > 0) {
                const int pos = atomicAdd(&s_n_eq, 1);
                if (pos < SELF_EQCAP) eqbuf[pos] = (uint32_t)i;
            }
        }
    }
    __syncthreads();
    int n_out = s_n_out;
    const int n_eq = s_n_eq;
    if (threadIdx.x == 0) overflow[q] = (need_eq > 0 && n_eq > SELF_EQCAP) ? 1 : 0;
    if (need_eq > 0) {
        if (n_eq > SELF_EQCAP) {
            // more ties at the threshold than the buffer holds: the row is redone by the radix kernel (flag above)
        } else {
            const int PE = next_pow2(n_eq > 1 ? n_eq : 1);
            for (int i = n_eq + threadIdx.x; i < PE; i += blockDim.x) eqbuf[i] = 0xFFFFFFFFu;
            __syncthreads();
            bitonic_sort_shared(eqbuf, PE);                   // ties: lowest indices first
            for (int t = threadIdx.x; t < need_eq; t += blockDim.x)
                keys[n_out + t] = ((KeyT)(uint32_t)(H - thr) << idx_bits) | (KeyT)eqbuf[t];
            n_out += need_eq;
        }
    }
    __syncthreads();
    bitonic_sort_shared(keys, P);
    const KeyT idx_mask = (((KeyT)1) << idx_bits) - 1;
    for (int t = threadIdx.x; t < kcand; t += blockDim.x) {
        int32_t oi = -1, oc = 0;
        if (t < n_out) {
            const KeyT k = keys[t];
            oi = (int32_t)(k & idx_mask);
            oc = H - (int32_t)(k >> idx_bits);
        }
        idx_out[q * kcand + t] = oi;
        cnt_out[q * kcand + t] = oc;
    }
    if (threadIdx.x == 0) nvalid_out[q] = n_out;
}

__global__ void fast_distance_kernel(const int32_t* __restrict__ idx_in, const int32_t* __restrict__ cnt_in,
                                     const int32_t* __restrict__ nvalid_in, int64_t nq, int kcand, int k, int H,
                                     int absolute, int32_t* __restrict__ idx_out, float* __restrict__ dist_out,
                                     int32_t* __restrict__ nvalid_out) {
    const int64_t total = nq * (int64_t)k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = e / k;
        const int t = (int)(e - q * k);
        const int nv = min(nvalid_in[q], k);
        int32_t oi = -1;
        float od = __int_as_float(0x7f800000);   // +inf padding
        if (t < nv) {
            oi = idx_in[q * kcand + t];
            const int c = cnt_in[q * kcand + t];
            od = absolute ? (float)c : 1.0f - __fdiv_rn((float)c, (float)H);   // rule 8, IEEE fp32
        }
        idx_out[e] = oi;
        dist_out[e] = od;
        if (t == 0) nvalid_out[q] = nv;
    }
}

static int bits_for(int64_t v) {   // bits needed to represent values in [0, v]
    int b = 1;
    while (((int64_t)1 << b) <= v) ++b;
    return b;
}

int launch_select_topk(const uint16_t* d_cnt, int64_t nq, int64_t ndb, int64_t ld_cnt, int H, int min_blocks,
                       int kcand, int32_t* d_idx, int32_t* d_cntsel, int32_t* d_nvalid, int* d_row_flags,
                       cudaStream_t stream) {
    if (nq <= 0) return 0;
    const int P = next_pow2(kcand);
    const int idx_bits = bits_for(ndb - 1 > 0 ? ndb - 1 : 1);
    const int cnt_bits = bits_for(H);
    const bool k32 = idx_bits + cnt_bits <= 32;
    const size_t smem = (size_t)P * (k32 ? 4 : 8);
    if (smem > 200 * 1024) return -1;   // caller reports SCHIC_ERR_UNSUPPORTED
    // single-histogram kernel when H is small enough and rows are 16-byte aligned; d_row_flags [nq] receives 1 for
    // the rows whose ties at the threshold did not fit its buffer, and the radix kernel then redoes exactly those
    const int nbins = (H + 256) & ~255;
    const size_t smem_full = smem + (size_t)nbins * 4 + (size_t)SELF_EQCAP * 4;
    const bool full = d_row_flags && H <= SELF_MAXH && (ld_cnt & 7) == 0 && ((uintptr_t)d_cnt & 15) == 0 &&
                      smem_full <= 200 * 1024;
    if (full) {
        if (k32) {
            cudaFuncSetAttribute(select_topk_fullhist_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full);
            select_topk_fullhist_kernel<uint32_t><<<(unsigned)nq, SEL_THREADS, smem_full, stream>>>(
                d_cnt, ndb, ld_cnt, H, min_blocks, kcand, P, idx_bits, d_idx, d_cntsel, d_nvalid, d_row_flags);
        } else {
            cudaFuncSetAttribute(select_topk_fullhist_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_full);
            select_topk_fullhist_kernel<uint64_t><<<(unsigned)nq, SEL_THREADS, smem_full, stream>>>(
                d_cnt, ndb, ld_cnt, H, min_blocks, kcand, P, idx_bits, d_idx, d_cntsel, d_nvalid, d_row_flags);
        }
    }
    const int* only = full ? d_row_flags : nullptr;
    if (k32) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(select_topk_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        select_topk_kernel<uint32_t><<<(unsigned)nq, SEL_THREADS, smem, stream>>>(
            d_cnt, ndb, ld_cnt, H, min_blocks, kcand, P, idx_bits, d_idx, d_cntsel, d_nvalid, only);
    } else {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(select_topk_kernel<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        select_topk_kernel<uint64_t><<<(unsigned)nq, SEL_THREADS, smem, stream>>>(
            d_cnt, ndb, ld_cnt, H, min_blocks, kcand, P, idx_bits, d_idx, d_cntsel, d_nvalid, only);
    }
    return full ? 2 : 1;
}

int launch_fast_distance(const int32_t* d_idx_in, const int32_t* d_cnt_in, const int32_t* d_nvalid_in, int64_t nq,
                         int kcand, int k, int H, int absolute, int32_t* d_idx_out, float* d_dist_out,
                         int32_t* d_nvalid_out, cudaStream_t stream) {
    if (nq <= 0) return 0;
    const int64_t total = nq * (int64_t)k;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    fast_distance_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_idx_in, d_cnt_in, d_nvalid_in, nq, kcand, k, H,
                                                                absolute, d_idx_out, d_dist_out, d_nvalid_out);
    return 1;
}

}  // namespace schic
