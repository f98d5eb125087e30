// k1_table.cu -- K1, shared-table formulation: MinHash signatures when many cells share features. sm_100a.
//
// Same contract as k1_signatures.cu (sig[c][j] = min over the cell's non-zero feature ids f of
// wang32((f+1)(j+1)) mod (2^31-1); call sites /root/reference/schicexplorer/scHicClusterMinHash.py:351,
// :353, :406), bit-identical output, different organisation of the work.
//
// The hash value of (feature f, hash function j) does not depend on the cell. A (cells x bins^2) Hi-C
// matrix re-uses every feature id in many cells (S10: 1.43e8 non-zeros, 3.8e6 distinct ids, x38), and the
// signature only ever keeps the SMALLEST value per (cell, j). So:
//   1. mark     present[f] = 1 for every id that occurs                                   (one pass over the CSR)
//   2. build    for every present id and every j evaluate the hash ONCE; keep (f, j) iff v < T, where
//               T = tau * 2^31 / mean_row_length -- about tau of a row's non-zeros fall below T for a given j.
//               Each CTA owns 256 consecutive ids and all H functions, collects its hits in shared memory and
//               publishes them as one contiguous run of 16-bit j's plus slot[f] = {start, count}.
//   3. lookup   a cell's signature is the minimum over its non-zeros of the hits stored for those ids:
//               one 8-byte slot read per non-zero (L2 resident), v re-derived from (f, j) by one evaluation,
//               atomicMin into a shared-memory signature.
//   4. fallback a (cell, j) whose minimum is >= T found no hit and still holds the sentinel M (real values are
//               < M): probability ~e^-tau, recomputed exactly over the cell's non-zeros.
// Exactness: every (f, j) with v < T is in the table, so if the true minimum of (cell, j) is < T the lookup sees
// it and everything else it sees is >= it; otherwise step 4 recomputes it. No value is approximated.
// Work: (distinct ids) x H evaluations instead of (non-zeros) x H, plus O(non-zeros) slot reads.
#include "common.cuh"

namespace schic {

constexpr int K1T_IDS = 256;          // feature ids per build CTA
constexpr int K1T_HITCAP = 3072;      // shared-memory hit list per build CTA (expected <= 2 hits per id)
constexpr int K1T_MAX_WARPS = 8;

__device__ __forceinline__ uint32_t k1t_mix(uint32_t a) {
    uint32_t b = a ^ (a >> 12);
    uint32_t c = b * 5u;
    uint32_t d = c ^ (c >> 4);
    uint32_t e = d * 2057u;
    return e ^ (e >> 16);
}
// exact wang32((f+1)(j+1)) mod M given fp1 = f+1 and c = 32767*(j+1)
__device__ __forceinline__ uint32_t k1t_value(uint32_t fp1, uint32_t c) {
    const uint32_t k6 = k1t_mix(fp1 * c - 1u);
    const uint32_t v = (k6 & kModulus) + (k6 >> 31);
    return v >= kModulus ? v - kModulus : v;
}
// width-64 HashSpec: the same formula in uint64 before the modulo (ids < 2^27 here, so f+1 has no high word)
__device__ __forceinline__ uint32_t k1t_value64(uint32_t fp1, uint32_t jp1) {
    uint64_t k = (uint64_t)fp1 * (uint64_t)jp1;
    k = ~k + (k << 15);
    k ^= k >> 12;
    k += k << 2;
    k ^= k >> 4;
    k *= 2057ull;
    k ^= k >> 16;
    uint64_t s = (k & kModulus) + ((k >> 31) & kModulus) + (k >> 62);    // fold 31-bit limbs (2^31 == 1 mod M)
    s = (s & kModulus) + (s >> 31);
    const uint32_t v = (uint32_t)s;
    return v >= kModulus ? v - kModulus : v;
}
template <bool W64>
__device__ __forceinline__ uint32_t k1t_value_w(uint32_t fp1, uint32_t j) {
    return W64 ? k1t_value64(fp1, j + 1u) : k1t_value(fp1, 32767u * (j + 1u));
}
// 15-bit prefix filter key (see k1_signatures.cu): v < best  =>  key <= threshold(best)
__device__ __forceinline__ uint32_t k1t_key(uint32_t a) {
    const uint32_t b = a ^ (a >> 12);
    const uint32_t c = b * 5u;
    const uint32_t d = c ^ (c >> 4);
    return d * 4114u + 0x20000u;
}
__host__ __device__ inline uint32_t k1t_threshold(uint32_t best) {
    const uint32_t hb = best >> 16;
    return hb >= 0x7FFEu ? 0xFFFFFFFFu : ((hb + 2u) << 17) - 1u;
}

// ---- 0. largest feature id (sizes the per-id arrays) ---------------------------------------------
__global__ void k1t_max_id_kernel(const uint32_t* __restrict__ feat, int64_t nnz, uint32_t* __restrict__ out_max) {
    uint32_t m = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) m = max(m, __ldg(feat + i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_max, m);
}

// ---- 1. mark ---------------------------------------------------------------------------------------
// counters[3] != 0 afterwards: some id was >= n_ids (only possible when n_ids comes from the caller's hint)
__global__ void k1t_mark_kernel(const uint32_t* __restrict__ feat, int64_t nnz, uint32_t n_ids,
                                uint8_t* __restrict__ present, uint32_t* __restrict__ counters) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const uint32_t f = __ldg(feat + i);
        if (f >= n_ids) { counters[3] = 1u; continue; }
        if (!present[f]) present[f] = 1;         // ids repeat ~40x: after the first touch this is a read that hits in L2
    }
}

// ---- 2. build --------------------------------------------------------------------------------------
// counters[0] = entries used, counters[1] = overflow flag (hit list or entry buffer too small)
template <int R, bool W64>
__global__ void __launch_bounds__(32 * K1T_MAX_WARPS)
k1t_build_kernel(const uint8_t* __restrict__ present, uint32_t n_ids, int H, int units, uint32_t T, uint32_t thrT,
                 uint2* __restrict__ slot, uint16_t* __restrict__ entries, uint32_t entries_cap,
                 uint32_t* __restrict__ counters) {
    __shared__ __align__(16) uint32_t plist[K1T_IDS + 4];      // local ids (0..255) that are present, padded to a multiple of 4
    __shared__ uint32_t hit[K1T_HITCAP];         // (local id << 16) | j
    __shared__ int cnt[K1T_IDS], excl[K1T_IDS], cur[K1T_IDS];
    __shared__ int s_np, s_nhit;
    __shared__ uint32_t s_base;

    const uint32_t f0 = blockIdx.x * (uint32_t)K1T_IDS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    if (threadIdx.x == 0) { s_np = 0; s_nhit = 0; }
    for (int i = threadIdx.x; i < K1T_IDS; i += blockDim.x) { cnt[i] = 0; cur[i] = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < K1T_IDS; i += blockDim.x) {
        const uint32_t f = f0 + i;
        if (f < n_ids && present[f]) plist[atomicAdd(&s_np, 1)] = (uint32_t)i;    // order is irrelevant
    }
    __syncthreads();
    const int np = s_np;
    if (np > 0) {
        const int np4 = (np + 3) & ~3;
        if (threadIdx.x < np4 - np) plist[np + threadIdx.x] = plist[np - 1];     // pad with a duplicate (never appended)
        __syncthreads();
        for (int u = warp; u < units; u += n_warps) {
            const int j0 = u * R * 32 + lane;
            uint32_t c[R];
#pragma unroll
            for (int r = 0; r < R; ++r) c[r] = 32767u * (uint32_t)(j0 + 32 * r + 1);
            if (W64) {                                              // no prefix filter for the 64-bit mix: plain evaluate + test
                for (int i = 0; i < np; ++i) {
                    const uint32_t li = plist[i];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int j = j0 + 32 * r;
                        if (j < H && k1t_value64(f0 + li + 1u, (uint32_t)j + 1u) < T) {
                            const int pos = atomicAdd(&s_nhit, 1);
                            if (pos < K1T_HITCAP) hit[pos] = (li << 16) | (uint32_t)j;
                        }
                    }
                }
                continue;
            }
            for (int i = 0; i < np4; i += 4) {
                const uint4 l4 = *reinterpret_cast<const uint4*>(plist + i);
                const uint32_t li[4] = {l4.x, l4.y, l4.z, l4.w};
                uint32_t fp1[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) fp1[q] = f0 + li[q] + 1u;
                uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t t0 = k1t_key(fp1[0] * c[r] - 1u), t1 = k1t_key(fp1[1] * c[r] - 1u);
                    const uint32_t t2 = k1t_key(fp1[2] * c[r] - 1u), t3 = k1t_key(fp1[3] * c[r] - 1u);
                    mn = min(mn, min(t0, t1));
                    mn = min(mn, min(t2, t3));
                }
                if (mn <= thrT) {                                   // rare: some (id, j) of this lane may be a hit
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int j = j0 + 32 * r;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (i + q < np && j < H && k1t_key(fp1[q] * c[r] - 1u) <= thrT &&
                                k1t_value(fp1[q], c[r]) < T) {
                                const int pos = atomicAdd(&s_nhit, 1);
                                if (pos < K1T_HITCAP) hit[pos] = (li[q] << 16) | (uint32_t)j;
                            }
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    const int nhit = s_nhit;
    if (nhit > K1T_HITCAP) {                                        // cannot happen for sane tau; reported, not hidden
        if (threadIdx.x == 0) atomicExch(&counters[1], 1u);
        return;
    }
    for (int h = threadIdx.x; h < nhit; h += blockDim.x) atomicAdd(&cnt[hit[h] >> 16], 1);
    __syncthreads();
    if (warp == 0) {                                                // exclusive scan of the 256 counts by one warp
        int loc[K1T_IDS / 32];
        int sum = 0;
#pragma unroll
        for (int q = 0; q < K1T_IDS / 32; ++q) { loc[q] = sum; sum += cnt[lane * (K1T_IDS / 32) + q]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int before = incl - sum;
#pragma unroll
        for (int q = 0; q < K1T_IDS / 32; ++q) excl[lane * (K1T_IDS / 32) + q] = before + loc[q];
        if (lane == 0) s_base = nhit > 0 ? atomicAdd(&counters[0], (uint32_t)nhit) : 0u;
    }
    __syncthreads();
    const uint32_t base = s_base;
    if (nhit > 0 && (uint64_t)base + (uint64_t)nhit > (uint64_t)entries_cap) {
        if (threadIdx.x == 0) atomicExch(&counters[1], 1u);
        return;
    }
    for (int h = threadIdx.x; h < nhit; h += blockDim.x) {
        const uint32_t li = hit[h] >> 16;
        entries[base + excl[li] + atomicAdd(&cur[li], 1)] = (uint16_t)(hit[h] & 0xFFFFu);
    }
    for (int i = threadIdx.x; i < K1T_IDS; i += blockDim.x)
        if (f0 + i < n_ids) slot[f0 + i] = make_uint2(base + (uint32_t)excl[i], (uint32_t)cnt[i]);
}

// ---- 3. lookup ---------------------------------------------------------------------------------------
struct K1TItem { int64_t begin; int32_t len; int32_t row; };
constexpr int K1T_PART = 8192;

// rows cut into parts of <= K1T_PART non-zeros; one entry per part (single CTA, order irrelevant here)
__global__ void __launch_bounds__(1024)
k1t_items_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, K1TItem* __restrict__ items, int* __restrict__ n_items) {
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
        const int64_t beg = indptr[r], len = indptr[r + 1] - beg;
        const int np = (int)((len + K1T_PART - 1) / K1T_PART);
        if (!np) continue;
        const int pos = atomicAdd(&s_n, np);
        for (int p = 0; p < np; ++p) {
            K1TItem it;
            it.begin = beg + (int64_t)p * K1T_PART;
            it.len = (int)min((int64_t)K1T_PART, len - (int64_t)p * K1T_PART);
            it.row = (int)r;
            items[pos + p] = it;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *n_items = s_n;
}

template <bool W64>
__global__ void __launch_bounds__(256)
k1t_lookup_kernel(const K1TItem* __restrict__ items, const int* __restrict__ n_items, const uint32_t* __restrict__ feat,
                  const uint2* __restrict__ slot, const uint16_t* __restrict__ entries, uint32_t* __restrict__ counters,
                  uint32_t n_ids, int H, uint32_t* __restrict__ sig) {
    extern __shared__ uint32_t sig_s[];          // [H]
    if ((int)blockIdx.x >= *n_items) return;
    if (counters[1] != 0u || counters[3] != 0u) return;   // table overflow / id beyond the table: everything goes to step 4
    const K1TItem it = items[blockIdx.x];
    for (int j = threadIdx.x; j < H; j += blockDim.x) sig_s[j] = kModulus;
    __syncthreads();
    const uint32_t* __restrict__ src = feat + it.begin;
    for (int i = threadIdx.x; i < it.len; i += blockDim.x) {
        const uint32_t f = __ldg(src + i);
        if (f >= n_ids) { counters[3] = 1u; continue; }      // beyond the declared range: reported by the host
        const uint2 s = __ldg(slot + f);
        for (uint32_t e = 0; e < s.y; ++e) {
            const uint32_t j = __ldg(entries + s.x + e);
            atomicMin(&sig_s[j], k1t_value_w<W64>(f + 1u, j));
        }
    }
    __syncthreads();
    uint32_t* __restrict__ sig_row = sig + (int64_t)it.row * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        const uint32_t v = sig_s[j];
        if (v < kModulus) atomicMin(sig_row + j, v);
    }
}

// ---- 4. fallback -------------------------------------------------------------------------------------
// One CTA per row: hash functions whose signature entry is still the sentinel are recomputed exactly. A warp
// takes one such j at a time and spreads the row's non-zeros over its lanes (coalesced loads, shuffle min):
// the common case is a long row with one or two failures, which a lane-per-j loop would walk serially.
template <bool W64>
__global__ void __launch_bounds__(128)
k1t_fallback_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat, int H,
                    uint32_t* __restrict__ sig) {
    extern __shared__ uint16_t fj[];             // [H] hash functions to redo
    __shared__ int s_n;
    const int64_t row = blockIdx.x;
    const int64_t b = indptr[row], e = indptr[row + 1];
    if (e <= b) return;                          // empty row: the sentinel IS the result
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    uint32_t* __restrict__ sig_row = sig + row * (int64_t)H;
    for (int j = threadIdx.x; j < H; j += blockDim.x)
        if (sig_row[j] == kModulus) fj[atomicAdd(&s_n, 1)] = (uint16_t)j;
    __syncthreads();
    const int n = s_n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int len = (int)(e - b);
    const uint32_t* __restrict__ src = feat + b;
    for (int t = warp; t < n; t += n_warps) {
        const uint32_t j = fj[t];
        uint32_t best = kModulus;
        int i = lane;
        for (; i + 96 < len; i += 128) {         // 4 independent loads per lane in flight
            const uint32_t f0 = __ldg(src + i), f1 = __ldg(src + i + 32), f2 = __ldg(src + i + 64), f3 = __ldg(src + i + 96);
            best = min(best, min(k1t_value_w<W64>(f0 + 1u, j), k1t_value_w<W64>(f1 + 1u, j)));
            best = min(best, min(k1t_value_w<W64>(f2 + 1u, j), k1t_value_w<W64>(f3 + 1u, j)));
        }
        for (; i < len; i += 32) best = min(best, k1t_value_w<W64>(__ldg(src + i) + 1u, j));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (lane == 0) sig_row[j] = best;
    }
}

__global__ void k1t_fill_kernel(uint32_t* __restrict__ p, int64_t n, uint32_t v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// ---- host side -----------------------------------------------------------------------------------------
// Threshold of the table: hash values below T = tau * 2^31 / mean_row are stored, so a (row, hash function) pair finds
// about tau stored candidates (none with probability e^-tau: those go to the fallback kernel) and an id stores
// H * tau / mean_row values. tau = 12 where rows are long; for short rows / many hash functions it is what
// K1T_MAX_PER_ID stored values per id allow.
constexpr double K1T_MAX_PER_ID = 8.0;
double k1t_tau(int64_t n_rows, int64_t nnz, int H) {
    const double tau = K1T_MAX_PER_ID * ((double)nnz / (double)n_rows) / (double)H;
    return tau > 12.0 ? 12.0 : tau;
}

// Table mode pays off when ids are shared: id range <= nnz / 2 (every id then occurs twice on average at
// least), rows long enough that tau >= 4 (under 2 % of the pairs fall back), and the per-id arrays stay small.
bool k1t_applicable(int64_t n_rows, int64_t nnz, uint32_t max_id, int H) {
    if (n_rows <= 0 || nnz <= 0 || H > 65535) return false;
    const double ids = (double)max_id + 1.0;
    if (ids > 134217728.0 || ids * 2.0 > (double)nnz) return false;
    return k1t_tau(n_rows, nnz, H) >= 4.0;
}

// expected stored values (every id of the range present: the prebuilt case) + 25 % + slack
size_t k1t_entries_capacity(int64_t n_rows, int64_t nnz, uint32_t max_id, int H) {
    const double ids = (double)max_id + 1.0;
    const double d = ids < (double)nnz ? ids : (double)nnz;
    double per_id = k1t_tau(n_rows, nnz, H) * (double)H / ((double)nnz / (double)n_rows);
    if (per_id < 2.0) per_id = 2.0;
    return (size_t)(1.25 * per_id * d) + 65536;
}

int64_t k1t_item_capacity(int64_t n_rows, int64_t nnz) { return n_rows + nnz / K1T_PART + 1; }

int launch_max_id(const uint32_t* d_feat, int64_t nnz, uint32_t* d_out_max, cudaStream_t stream) {
    if (cudaMemsetAsync(d_out_max, 0, sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (nnz <= 0) return 0;
    k1t_max_id_kernel<<<148 * 8, 256, 0, stream>>>(d_feat, nnz, d_out_max);
    return 1;
}

static int pick_R_table(int slabs) {
    int best = 5, best_waste = 1 << 30;
    for (int R = 6; R >= 4; --R) {
        const int waste = ((slabs + R - 1) / R) * R - slabs;
        if (waste < best_waste) { best_waste = waste; best = R; }
    }
    return best;
}

// d_present [max_id+1] bytes, d_slot [max_id+1], d_entries [entries_cap], d_counters [4] uint32 ([3]: id >= max_id+1
// seen, sticky), d_items/d_n_items:
// k1t_item_capacity entries + 1 int.
// phase 0: everything. phase 1: build the table only, for EVERY id <= max_id (no look at the data: it can run while
// the ids are still being uploaded; needs only n_rows / nnz for the threshold). phase 2: lookup + fallback against
// a table built by phase 1 with the same n_rows / nnz.
int launch_signatures_table(int width, const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows, int64_t nnz,
                              int64_t indptr0, uint32_t max_id, int H, uint32_t* d_sig, uint8_t* d_present,
                              uint2* d_slot, uint16_t* d_entries, size_t entries_cap, uint32_t* d_counters,
                              void* d_items, int phase, cudaStream_t stream) {
    if (n_rows <= 0) return 0;
    if (nnz <= 0) {
        if (phase != 1) k1t_fill_kernel<<<148 * 8, 256, 0, stream>>>(d_sig, n_rows * (int64_t)H, kModulus);
        return phase != 1 ? 1 : 0;
    }
    const uint32_t n_ids = max_id + 1u;
    const double mean_row = (double)nnz / (double)n_rows;
    const double tau = k1t_tau(n_rows, nnz, H);
    double Td = tau * 2147483647.0 / mean_row;
    if (Td > 2147483646.0) Td = 2147483646.0;
    const uint32_t T = (uint32_t)Td;
    const uint32_t thrT = k1t_threshold(T);
    int nl = 0;
    if (phase != 1) { k1t_fill_kernel<<<148 * 8, 256, 0, stream>>>(d_sig, n_rows * (int64_t)H, kModulus); ++nl; }
    const int slabs = (H + 31) / 32;
    if (phase != 2) {
    if (phase == 1) {
        cudaMemsetAsync(d_present, 1, (size_t)n_ids, stream);
    } else {
        cudaMemsetAsync(d_present, 0, (size_t)n_ids, stream);
    }
    cudaMemsetAsync(d_counters, 0, 2 * sizeof(uint32_t), stream);      // [3] (range violation) is sticky: cleared by the owner
    if (phase == 0) { k1t_mark_kernel<<<148 * 8, 256, 0, stream>>>(d_feat + indptr0, nnz, n_ids, d_present, d_counters); ++nl; }
    const int R = pick_R_table(slabs);
    const int units = (slabs + R - 1) / R;
    int n_warps = units < K1T_MAX_WARPS ? units : K1T_MAX_WARPS;
    for (int w = K1T_MAX_WARPS; w >= 4; --w)
        if (units % w == 0) { n_warps = w; break; }
    const unsigned build_grid = (n_ids + K1T_IDS - 1) / K1T_IDS;
#define K1T_BUILD(RR)                                                                                                  \
    do {                                                                                                               \
        if (width == 64)                                                                                               \
            k1t_build_kernel<RR, true><<<build_grid, 32 * n_warps, 0, stream>>>(d_present, n_ids, H, units, T, thrT,    \
                                                                                d_slot, d_entries, (uint32_t)entries_cap, d_counters); \
        else                                                                                                           \
            k1t_build_kernel<RR, false><<<build_grid, 32 * n_warps, 0, stream>>>(d_present, n_ids, H, units, T, thrT,   \
                                                                                 d_slot, d_entries, (uint32_t)entries_cap, d_counters); \
    } while (0)
    switch (R) {
        case 4: K1T_BUILD(4); break;
        case 5: K1T_BUILD(5); break;
        default: K1T_BUILD(6); break;
    }
#undef K1T_BUILD
    ++nl;
    }
    if (phase == 1) return nl;
    const int64_t cap = k1t_item_capacity(n_rows, nnz);
    K1TItem* items = reinterpret_cast<K1TItem*>(d_items);
    int* n_items = reinterpret_cast<int*>(items + cap);
    k1t_items_kernel<<<1, 1024, 0, stream>>>(d_indptr, n_rows, items, n_items); ++nl;
    if (width == 64) {
        cudaFuncSetAttribute(k1t_lookup_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, H * 4);
        k1t_lookup_kernel<true><<<(unsigned)cap, 256, (size_t)H * 4, stream>>>(items, n_items, d_feat, d_slot, d_entries, d_counters, n_ids, H, d_sig); ++nl;
        cudaFuncSetAttribute(k1t_fallback_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, H * 2);
        k1t_fallback_kernel<true><<<(unsigned)n_rows, 128, (size_t)H * 2, stream>>>(d_indptr, d_feat, H, d_sig); ++nl;
    } else {
        cudaFuncSetAttribute(k1t_lookup_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, H * 4);
        k1t_lookup_kernel<false><<<(unsigned)cap, 256, (size_t)H * 4, stream>>>(items, n_items, d_feat, d_slot, d_entries, d_counters, n_ids, H, d_sig); ++nl;
        cudaFuncSetAttribute(k1t_fallback_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, H * 2);
        k1t_fallback_kernel<false><<<(unsigned)n_rows, 128, (size_t)H * 2, stream>>>(d_indptr, d_feat, H, d_sig); ++nl;
    }
    return nl;
}

size_t k1t_items_bytes(int64_t n_rows, int64_t nnz) { return (size_t)k1t_item_capacity(n_rows, nnz) * sizeof(K1TItem) + 16; }

}  // namespace schic
