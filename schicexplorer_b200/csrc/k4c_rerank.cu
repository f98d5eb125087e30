// k4c_rerank.cu -- K4 for integer contact counts: packed candidate stream + byte table, sm_100a.
//
// Same contract as k4_rerank.cu (the fast=False branch of MinHash.kneighbors, selected by `-em`,
// /root/reference/schicexplorer/scHicClusterMinHash.py:66-68 -> :348/:403; SURVEY 8(c) rule 9): for every query
// cell and each of its candidates the euclidean distance of the two raw sparse rows, (distance asc, index asc), keep k.
//
// Hi-C values are contact counts -- small non-negative integers. For such data the rerank is organised differently
// from the generic float kernel (k4_rerank_window_kernel, 2 x 4-byte loads and ~23 issued instructions per candidate
// non-zero, 58.7 GB of DRAM traffic per S10 build):
//   * per fit ONE pass over the CSR (k4c_prep_kernel) produces, besides the fp64 row norms and the window directory,
//     a PACKED STREAM: one 32-bit word per non-zero = (id mod 2^16) << 16 | count. Half the bytes and half the loads
//     of (id, float value); also what a multi-GPU run exchanges instead of the two CSR arrays.
//   * the id axis is cut into windows of 2^16 ids. For one window the query row is a DIRECT BYTE TABLE in shared
//     memory (64 KB): tab[id mod 2^16] = count (0 = absent). A probe is one LDS.U8 and one IDP.2A -- a miss
//     multiplies by zero, so there is no bit test, no rank/popcount, no second load and no predicate.
//     Counts >= 255 in a query row are stored as 255 and resolved exactly by a (rare) binary search in the query row;
//     the build phase knows whether a window has any, so windows without them run a loop that never checks.
//   * x.y is accumulated in integers (8 products < 2^26 in a 32-bit register, then 64 bit), ||x-y||^2 =
//     ||x||^2 + ||y||^2 - 2 x.y is exact, so there is no cancellation case and no second exact pass.
//   * the grid is WINDOW-GROUP MAJOR: CTA = (group of consecutive windows, query), groups slow-varying, so the 148
//     resident CTAs stream the same ~1/G slice of the database at the same time and it is served by the 126 MB L2
//     instead of HBM. Partial dot products go to a global uint64 accumulator (one atomic per candidate and CTA);
//     k4c_finalize_kernel turns them into distances and sorts.
// Data that is not small integer counts (non-integer, negative, > 65535, or so many >= 255 that the escape path
// would dominate) is detected by the prep pass in a device flag; these kernels then exit at once and the generic
// float kernels of k4_rerank.cu, launched behind them with the same flag, do the work. No host round trip decides.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sort.cuh"

namespace schic {

// must match k4_rerank.cu
struct K4Chunk {
    int cand_stride, cand_off, out_stride, out_off, write_nvalid, n_all;
};

constexpr int K4C_WIN_BITS = SCHIC_K4C_WIN_BITS;
constexpr int K4C_TAB = 1 << K4C_WIN_BITS;                    // table bytes = ids per window
constexpr int K4C_CNT_BITS = 32 - K4C_WIN_BITS;               // 16
constexpr uint32_t K4C_CNT_MASK = (1u << K4C_CNT_BITS) - 1u;  // largest packable count (65535)
// count x table byte, accumulated. With the 16/16 split one IDP.2A does it (the id half of the word meets a zero byte).
__device__ __forceinline__ uint32_t k4c_mac(uint32_t word, uint32_t table_byte, uint32_t acc) {
    if (K4C_WIN_BITS == 16) return __dp2a_lo(word, table_byte, acc);
    return acc + table_byte * (word & K4C_CNT_MASK);
}
// 64 KB table (2^16-id windows): 512 threads, two CTAs per SM. Experiment build -DSCHIC_K4C_WIN_BITS=17: 128 KB table,
// 1024 threads, one CTA per SM, counts up to 32767.
#ifndef K4C_THREADS_DEF
#define K4C_THREADS_DEF (SCHIC_K4C_WIN_BITS == 16 ? 512 : 1024)
#endif
constexpr int K4C_THREADS = K4C_THREADS_DEF;                  // (a multiple of 32)
#ifndef K4C_CTAS_DEF
#define K4C_CTAS_DEF (SCHIC_K4C_WIN_BITS == 16 ? 2 : 1)
#endif
constexpr int K4C_CTAS = K4C_CTAS_DEF;                        // CTAs per SM the kernel is built for (register budget)

__device__ __forceinline__ bool k4c_usable(const uint32_t* __restrict__ flags, uint32_t esc_limit) {
    return flags[0] == 0u && flags[1] <= esc_limit;
}

int k4c_window_count(uint32_t max_feature) { return (int)(max_feature >> K4C_WIN_BITS) + 1; }

// ---- per-fit preparation: norms, window directory, packed stream, flags -- one pass over ids + values -------------
// dir[row][w] = offset (from the row start) of the first non-zero with (id >> 17) >= w, w = 0..n_windows.
// flags[0] |= 1: a value is not an integer in [0, 65535]; flags[1] += number of values >= 255; flags[2] = 1: an id
// lies beyond the last window (wrong feature-range hint; such ids are clamped into a window no slice covers).
// packed may be nullptr (directory + norms only); dir may be nullptr (norms + flags only).
__global__ void __launch_bounds__(256)
k4c_prep_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat, const float* __restrict__ val,
                int64_t n_rows, int n_windows, double* __restrict__ norm2, uint32_t* __restrict__ dir,
                uint32_t* __restrict__ packed, uint32_t* __restrict__ flags) {
    __shared__ double s_acc[8];
    const int64_t row = blockIdx.x;
    if (row >= n_rows) return;
    const int64_t b = indptr[row], e = indptr[row + 1];
    uint32_t* __restrict__ d = dir ? dir + row * (int64_t)(n_windows + 1) : nullptr;
    double acc = 0.0;
    uint32_t bad = 0u, big = 0u, oor = 0u;
#pragma unroll 4
    for (int64_t i = b + threadIdx.x; i < e; i += 256) {
        const uint32_t f = __ldg(feat + i);
        const float v = __ldg(val + i);
        acc += (double)v * (double)v;
        const bool ok = v >= 0.f && v <= (float)K4C_CNT_MASK && v == truncf(v);      // NaN fails every comparison
        bad |= ok ? 0u : 1u;
        const uint32_t c = ok ? (uint32_t)v : 0u;
        big += c >= 255u ? 1u : 0u;
        if (packed) packed[i] = ((f & (uint32_t)(K4C_TAB - 1)) << K4C_CNT_BITS) | c;
        if (d) {
            int w = (int)(f >> K4C_WIN_BITS);
            if (w >= n_windows) { oor = 1u; w = n_windows; }
            const int wp = i == b ? -1 : min((int)(__ldg(feat + i - 1) >> K4C_WIN_BITS), n_windows);
            for (int x = wp + 1; x <= w; ++x) d[x] = (uint32_t)(i - b);
        }
    }
    if (d) {
        const int wl = e > b ? min((int)(__ldg(feat + e - 1) >> K4C_WIN_BITS), n_windows) : -1;
        for (int x = wl + 1 + threadIdx.x; x <= n_windows; x += 256) d[x] = (uint32_t)(e - b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        big += __shfl_xor_sync(0xffffffffu, big, o);
    }
    bad = __any_sync(0xffffffffu, bad) ? 1u : 0u;
    oor = __any_sync(0xffffffffu, oor) ? 1u : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_acc[warp] = acc;
        if (bad) atomicOr(&flags[0], 1u);
        if (big) atomicAdd(&flags[1], big);
        if (oor) flags[2] = 1u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_acc[w];      // fixed order: the norm is reproducible
        norm2[row] = t;
    }
}

int launch_rerank_prep(const int64_t* d_indptr, const uint32_t* d_feat, const float* d_val, int64_t n_rows, int n_windows,
                       double* d_norm2, uint32_t* d_dir, uint32_t* d_packed, uint32_t* d_flags, cudaStream_t stream) {
    if (n_rows <= 0) return 0;
    k4c_prep_kernel<<<(unsigned)n_rows, 256, 0, stream>>>(d_indptr, d_feat, d_val, n_rows, n_windows, d_norm2, d_dir,
                                                          d_packed, d_flags);
    return 1;
}

// exact count of query key t (window-relative id) whose table byte saturated at 255: binary search in the query's
// window slice of the packed stream (keys ascend with the ids)
__device__ __noinline__ uint32_t k4c_escape_value(const uint32_t* __restrict__ qp, int qn, uint32_t t) {
    int lo = 0, hi = qn;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((__ldg(qp + mid) >> K4C_CNT_BITS) < t) lo = mid + 1; else hi = mid;
    }
    return lo < qn ? (__ldg(qp + lo) & K4C_CNT_MASK) : 0u;
}

// Candidate order of every query for the rerank kernel: columns of the launch's candidate slice sorted by row length
// (descending, then by column). The rerank kernel gives each slice to 8 lanes and runs four slices per warp in lock
// step, so neighbours in this order -- rows of similar length, hence window slices of similar length -- share a warp.
// lane_map != nullptr (kcand <= 128, 256 threads): additionally the lane assignment of the rerank kernel's CTA. Row
// lengths of one query's candidates differ by an order of magnitude (log-normal depth), and with one fixed-size lane
// group per candidate every window waits for the longest slice. Here candidate i gets gs_i lanes, a power of two
// proportional to its row length (4..32 -- a slice's <= 3 head / tail elements need 3 lanes --, sum <= K4C_THREADS),
// groups packed in descending size (hence aligned): lane_map[q][lane] = (sorted position << 8) | gs, 0xFFFF = spare lane.
__global__ void __launch_bounds__(256)
k4c_order_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ cand_idx,
                 const int32_t* __restrict__ cand_nvalid, int kcand, int P, K4Chunk ck, const uint32_t* __restrict__ flags,
                 uint32_t esc_limit, int32_t* __restrict__ order, uint16_t* __restrict__ lane_map) {
    if (!k4c_usable(flags, esc_limit)) return;
    extern __shared__ unsigned long long okeys[];
    __shared__ int s_scan[40];
    const int64_t q = blockIdx.x;
    const int ncand = max(0, min((cand_idx ? cand_nvalid[q] : ck.n_all) - ck.cand_off, kcand));
    for (int c = threadIdx.x; c < P; c += blockDim.x) {
        unsigned long long key = ~0ull;
        if (c < ncand) {
            const int32_t d = cand_idx ? cand_idx[q * ck.cand_stride + ck.cand_off + c] : ck.cand_off + c;
            const uint32_t len = (uint32_t)(indptr[d + 1] - indptr[d]);
            key = ((unsigned long long)(~len) << 32) | (unsigned long long)(uint32_t)c;
        }
        okeys[c] = key;
    }
    bitonic_sort_shared(okeys, P);
    for (int c = threadIdx.x; c < kcand; c += blockDim.x) order[q * (int64_t)kcand + c] = c < ncand ? (int32_t)(uint32_t)okeys[c] : -1;
    if (!lane_map) return;
    // proportional lane groups (blockDim.x == 256 >= P: one sorted candidate per thread)
    const int i = threadIdx.x;
    const uint32_t len = i < ncand ? ~(uint32_t)(okeys[i] >> 32) : 0u;
    int total = 0;
    block_exclusive_scan((int)min(len, 0x3FFFFFu), s_scan, &total);
    uint32_t t = max(1u, ((uint32_t)total + K4C_THREADS - 1) / K4C_THREADS);       // non-zeros per lane to aim for
    int gs = 0, start = 0;
    for (int it = 0; it < 24; ++it) {
        gs = 0;
        if (i < ncand) {
            const uint32_t want = (len + t - 1) / t;
            gs = 4;
            while (gs < 32 && (uint32_t)gs < want) gs <<= 1;
        }
        int sum = 0;
        start = block_exclusive_scan(gs, s_scan, &sum);
        if (sum <= K4C_THREADS) break;
        t += t / 4 + 1;
        if (it == 23) { gs = i < ncand ? 4 : 0; start = 4 * i; }    // (ncand <= 128: groups of 4 always fit)
    }
    uint16_t* __restrict__ m = lane_map + q * (int64_t)K4C_THREADS;
    for (int l = threadIdx.x; l < K4C_THREADS; l += blockDim.x) m[l] = 0xFFFFu;
    __syncthreads();
    for (int l = 0; l < gs; ++l) m[start + l] = (uint16_t)((i << 8) | gs);
}

// One CTA = (window group g, query q); blockIdx.x = g * nq + q. dotacc [nq, kcand] uint64, zeroed by the caller.
// 512 threads and a 64 KB table, so that K4C_CTAS CTAs share an SM: while one waits at a barrier or for its loads, the
// others issue. Per window three barriers (un-build | build | stream). A window's slice of one candidate (~120 non-zeros
// at S10) is far too short to amortise a warp-wide set-up and reduction, so the unit of work is 8 lanes per slice, four
// slices per warp, 64 per CTA, statically assigned in the length-sorted candidate order: no hand-out atomics, a 3-step
// reduction. Every lane group walks its own sequence of BATCHES (NB x 16 bytes per lane; a typical slice is one to
// four batches) -- slice after slice, window after window -- and always holds the next batch of that sequence in
// registers while the current one is probed: loads cross the barriers (they do not depend on the table), as do the
// query's keys of the next window, which later un-build the table from the same registers.
// A probe is LEA.HI (table address), LDS.U8, IDP.2A (count x table byte accumulated; the id half of the word meets a
// zero byte): three instructions per candidate non-zero.
template <int LPS>     // lanes per slice: 8 or 4 (fixed groups, candidates dealt round-robin), 0 = per-candidate group sizes from lane_map
__global__ void __launch_bounds__(K4C_THREADS, K4C_CTAS)
k4c_rerank_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ packed,
                  const uint32_t* __restrict__ dir, int n_windows, int wg, int64_t q_row0, int64_t nq,
                  const int32_t* __restrict__ cand_idx, const int32_t* __restrict__ cand_nvalid,
                  const int32_t* __restrict__ order, const uint16_t* __restrict__ lane_map, int kcand, K4Chunk ck,
                  const uint32_t* __restrict__ flags, uint32_t esc_limit, unsigned long long* __restrict__ dotacc) {
    if (!k4c_usable(flags, esc_limit)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t* tab = smem_raw;                                                            // [K4C_TAB]
    unsigned long long* dotc = reinterpret_cast<unsigned long long*>(smem_raw + K4C_TAB);   // [kcand] by sorted position
    uint32_t* sl = reinterpret_cast<uint32_t*>(dotc + kcand);                           // [kcand][wg + 1] absolute positions
    uint32_t* qdir = sl + (size_t)kcand * (wg + 1);                                     // [wg + 1]
    __shared__ int s_esc;

    const int64_t g = blockIdx.x / nq;
    const int64_t q = blockIdx.x - g * nq;
    const int64_t qrow = q_row0 + q;
    const int w0 = (int)g * wg;
    const int nwin = min(wg, n_windows - w0);
    const int64_t nw1 = (int64_t)n_windows + 1;
    const int wg1 = wg + 1;
    const int lane = threadIdx.x & 31;
#ifndef K4C_NB_DEF
#define K4C_NB_DEF 2          // measured at S10 (groups of 4 lanes): 1 -> 16.8 ms, 2 -> 14.5 ms, 3 -> 17.2 ms, 4 -> 15.1 ms
#endif
    constexpr int NB = K4C_NB_DEF;                         // 16-byte words per lane and batch
    const int ncand = max(0, min((cand_idx ? cand_nvalid[q] : ck.n_all) - ck.cand_off, kcand));
    // this lane's group: gs lanes (a power of two, aligned), ql = rank in it, first_i = its first sorted candidate
    // position (-1: spare lane), stride_i = step to its next candidate of the same window
    int gs = LPS, first_i, stride_i;
    if (LPS > 0) {
        first_i = threadIdx.x / (LPS > 0 ? LPS : 1);
        stride_i = K4C_THREADS / (LPS > 0 ? LPS : 1);
        if (first_i >= ncand) first_i = -1;
    } else {
        const uint32_t m = __ldg(lane_map + q * (int64_t)K4C_THREADS + threadIdx.x);
        gs = m == 0xFFFFu ? 4 : (int)(m & 0xFFu);
        first_i = m == 0xFFFFu ? -1 : (int)(m >> 8);
        stride_i = kcand + 1;                              // one candidate per group: the next one is in the next window
    }
    const int ql = lane & (gs - 1);
    const uint32_t BATCH = (uint32_t)gs * NB * 4u;         // non-zeros per batch of the group
    const uint32_t qmask = (gs == 32 ? 0xFFFFFFFFu : ((1u << gs) - 1u)) << (lane & ~(gs - 1));

    for (int w = threadIdx.x; w <= nwin; w += K4C_THREADS) qdir[w] = __ldg(dir + qrow * nw1 + w0 + w);
    if (threadIdx.x == 0) s_esc = 0;
    __syncthreads();
    if (qdir[0] == qdir[nwin] || ncand == 0) return;       // the query row is empty in this whole group (uniform)

    const int64_t xb = indptr[qrow];
    // first non-empty window of the query, and its keys (two per thread in registers; longer windows loop)
    auto next_window = [&](int j) { while (j < nwin && qdir[j + 1] == qdir[j]) ++j; return j; };
    uint32_t qn0 = 0u, qn1 = 0u;
    auto prefetch_keys = [&](int j) {
        if (j >= nwin) return;
        const uint32_t i0 = qdir[j] + threadIdx.x, e = qdir[j + 1];
        if (i0 < e) qn0 = __ldg(packed + xb + i0);
        if (i0 + K4C_THREADS < e) qn1 = __ldg(packed + xb + i0 + K4C_THREADS);
    };
    int j = next_window(0);
    prefetch_keys(j);

    for (int c = threadIdx.x; c < ncand; c += K4C_THREADS) dotc[c] = 0ull;
    {
        // absolute stream positions of every (sorted candidate, window border) of this group (total nnz < 2^32)
        const int per = nwin + 1, total = ncand * per;
        for (int x = threadIdx.x; x < total; x += K4C_THREADS) {
            const int i = x / per, jj = x - i * per;
            const int c = __ldg(order + q * (int64_t)kcand + i);
            const int32_t d = cand_idx ? cand_idx[q * ck.cand_stride + ck.cand_off + c] : ck.cand_off + c;
            sl[i * wg1 + jj] = (uint32_t)(indptr[d] + (int64_t)__ldg(dir + (int64_t)d * nw1 + w0 + jj));
        }
    }
    for (int i = threadIdx.x; i < K4C_TAB / 16; i += K4C_THREADS)
        reinterpret_cast<uint4*>(tab)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();                                       // sl complete

    // ---- the lane group's batch sequence -----------------------------------------------------------------------
    // A slice [ys, ye) is read as its 16-byte aligned interior [A, B) with 128-bit loads (lane ql takes words ql, ql + 8,
    // ... of every batch; words beyond B are zero and contribute nothing) plus <= 3 head and <= 3 tail elements: one
    // 32-bit load on lanes 0-2 / 4-6 of the group, probed with the slice's first batch.
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    int it_i = first_i - stride_i, it_j = j;               // slice the iterator reads (advanced before use)
    uint32_t it_p = 0u, it_B = 0u;                         // next batch position / interior end of that slice
    uint4 h[NB];                                           // look-ahead batch
    uint32_t h_sx = 0u, h_tx = 0u;                         // its head / tail boundary element (first batch of a slice, else 0)
    uint32_t h_sx2 = 0u, h_tx2 = 0u;                       // groups of 2 lanes: the third boundary element goes to lane 0
    int h_i = 0, h_j = nwin;                               // its slice; h_j == nwin: the sequence is exhausted
    bool h_last = false;
    auto advance = [&]() {
        h_sx = 0u; h_tx = 0u; h_sx2 = 0u; h_tx2 = 0u;
        if (it_p >= it_B) {                                // open the next non-empty slice of this lane group
            uint32_t ys = 0u, ye = 0u;
            for (;;) {
                it_i += stride_i;
                if (it_i >= ncand) { it_i = first_i; it_j = next_window(it_j + 1); }
                if (it_j >= nwin) { h_j = nwin; return; }
                ys = sl[it_i * wg1 + it_j]; ye = sl[it_i * wg1 + it_j + 1];
                if (ye > ys) break;
            }
            it_p = min((ys + 3u) & ~3u, ye);
            it_B = max(it_p, ye & ~3u);
            if (ql < (int)(it_p - ys)) h_sx = __ldg(packed + ys + ql);           // <= 3 elements before the interior
            if (ql < (int)(ye - it_B)) h_tx = __ldg(packed + it_B + ql);         // <= 3 elements behind it
            if (LPS == 2) {
                if (ql + 2 < (int)(it_p - ys)) h_sx2 = __ldg(packed + ys + ql + 2);
                if (ql + 2 < (int)(ye - it_B)) h_tx2 = __ldg(packed + it_B + ql + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const uint32_t p = it_p + 4u * ql + 4u * gs * u;
            h[u] = p < it_B ? __ldg(reinterpret_cast<const uint4*>(packed + p)) : zero4;
        }
        h_i = it_i; h_j = it_j;
        it_p += BATCH;
        h_last = it_p >= it_B;
    };
    if (first_i >= 0) advance();

    unsigned long long acc = 0ull;
    while (j < nwin) {
        const uint32_t k0 = qdir[j], k1 = qdir[j + 1];
        const uint32_t qc0 = qn0, qc1 = qn1;               // this window's keys (prefetched)
        __syncthreads();                                   // table clear / previous un-build done
        // ---- build: tab[id mod 2^16] = min(count, 255) ------------------------------------------------------
        bool any_big = false;
        {
            const uint32_t i0 = k0 + threadIdx.x;
            if (i0 < k1) { const uint32_t v = qc0 & K4C_CNT_MASK; tab[qc0 >> K4C_CNT_BITS] = (uint8_t)min(v, 255u); any_big |= v >= 255u; }
            if (i0 + K4C_THREADS < k1) { const uint32_t v = qc1 & K4C_CNT_MASK; tab[qc1 >> K4C_CNT_BITS] = (uint8_t)min(v, 255u); any_big |= v >= 255u; }
            for (uint32_t i = i0 + 2 * K4C_THREADS; i < k1; i += K4C_THREADS) {
                const uint32_t p = __ldg(packed + xb + i);
                const uint32_t v = p & K4C_CNT_MASK;
                tab[p >> K4C_CNT_BITS] = (uint8_t)min(v, 255u);
                any_big |= v >= 255u;
            }
        }
        if (any_big) s_esc = j + 1;                        // (j + 1 never repeats within a CTA: no reset needed)
        const int jn = next_window(j + 1);
        prefetch_keys(jn);                                 // consumed at the top of the next iteration
        __syncthreads();
        // ---- stream: this lane group's batches of window j ------------------------------------------------------------
        {
            const bool esc = s_esc == j + 1;
            const uint32_t* __restrict__ qp = packed + xb + k0;
            const int qn = (int)(k1 - k0);
            while (h_j == j) {
                uint4 v[NB];
#pragma unroll
                for (int u = 0; u < NB; ++u) v[u] = h[u];
                const uint32_t sx = h_sx, tx = h_tx, sx2 = h_sx2, tx2 = h_tx2;
                const int cur_i = h_i;
                const bool last = h_last;
                advance();                                 // next batch in flight while this one is probed
                if (!esc) {
                    uint32_t a32 = k4c_mac(sx, (uint32_t)tab[sx >> K4C_CNT_BITS], 0u);   // <= 18 products < 2^24 each
                    a32 = k4c_mac(tx, (uint32_t)tab[tx >> K4C_CNT_BITS], a32);
                    if (LPS == 2) {
                        a32 = k4c_mac(sx2, (uint32_t)tab[sx2 >> K4C_CNT_BITS], a32);
                        a32 = k4c_mac(tx2, (uint32_t)tab[tx2 >> K4C_CNT_BITS], a32);
                    }
#pragma unroll
                    for (int u = 0; u < NB; ++u) {
                        a32 = k4c_mac(v[u].x, (uint32_t)tab[v[u].x >> K4C_CNT_BITS], a32);
                        a32 = k4c_mac(v[u].y, (uint32_t)tab[v[u].y >> K4C_CNT_BITS], a32);
                        a32 = k4c_mac(v[u].z, (uint32_t)tab[v[u].z >> K4C_CNT_BITS], a32);
                        a32 = k4c_mac(v[u].w, (uint32_t)tab[v[u].w >> K4C_CNT_BITS], a32);
                    }
                    acc += a32;
                } else {
                    // this window of the query holds counts >= 255: table bytes of 255 are looked up exactly
                    auto probe = [&](uint32_t w) {
                        const uint32_t t = w >> K4C_CNT_BITS, cnt = w & K4C_CNT_MASK;
                        const uint32_t tv = tab[t];
                        acc += (unsigned long long)(tv * cnt);
                        if (tv == 255u && cnt != 0u)
                            acc += (unsigned long long)(k4c_escape_value(qp, qn, t) - 255u) * (unsigned long long)cnt;
                    };
                    probe(sx);
                    probe(tx);
                    if (LPS == 2) { probe(sx2); probe(tx2); }
#pragma unroll
                    for (int u = 0; u < NB; ++u) { probe(v[u].x); probe(v[u].y); probe(v[u].z); probe(v[u].w); }
                }
                if (last) {
                    for (int o = 1; o < gs; o <<= 1) acc += __shfl_xor_sync(qmask, acc, o);
                    if (ql == 0) dotc[cur_i] += acc;       // only this lane group touches position cur_i
                    acc = 0ull;
                }
            }
        }
        // ---- un-build: zero exactly the bytes this window set ----------------------------------------------------
        __syncthreads();
        {
            const uint32_t i0 = k0 + threadIdx.x;
            if (i0 < k1) tab[qc0 >> K4C_CNT_BITS] = 0;
            if (i0 + K4C_THREADS < k1) tab[qc1 >> K4C_CNT_BITS] = 0;
            for (uint32_t i = i0 + 2 * K4C_THREADS; i < k1; i += K4C_THREADS) tab[__ldg(packed + xb + i) >> K4C_CNT_BITS] = 0;
        }
        j = jn;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncand; i += K4C_THREADS)
        if (dotc[i] != 0ull) atomicAdd(&dotacc[q * (int64_t)kcand + __ldg(order + q * (int64_t)kcand + i)], dotc[i]);
}

// ---- direct variant: no batch iterator ----------------------------------------------------------------------------
// Same CTA = (window group, query), same table, same shared-memory layout and the same three barriers per window as
// k4c_rerank_kernel, but the streaming phase is a plain loop: lane group gi (LPS lanes) owns the candidates at sorted
// positions gi, gi + NG, ...; in one window it reads such a candidate's slice as ALIGNED 16-byte chunks, a round of
// U chunks per lane issued back to back (LPS * U * 4 non-zeros; one round covers the typical slice), then probes them.
// No head / tail special cases inside the loop: an aligned chunk that straddles the slice's start or end carries <= 3
// foreign non-zeros (of the neighbouring window or row); those are probed like the rest and their products are
// subtracted again -- <= 3 + 3 single-word loads per slice on lanes 0-2 of the group, exact in modular arithmetic
// (the stream is allocated with 4 words of slack behind its end). Round 0 of the group's first slice of the NEXT
// window is requested right after the probes of this window, so those loads are in flight across un-build, the
// barriers and the next build: the latency hiding of the batch iterator without its ~280 instructions of per-window
// bookkeeping per warp (ncu r02n: 76 % of the 8.9e9 issued instructions of k4c_rerank_kernel were not probes).
// T = threads per CTA: NG = T / LPS lane groups should cover the candidate list (warps without a candidate still pay
// for the three barriers per window: 100 candidates x 4 lanes on 416 threads 10.4 ms, on 512 threads 11.2 ms at S10).
template <int LPS, int U, int T>
__global__ void __launch_bounds__(T, K4C_CTAS)
k4c_rerank_direct_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ packed,
                         const uint32_t* __restrict__ dir, int n_windows, int wg, int64_t q_row0, int64_t nq,
                         const int32_t* __restrict__ cand_idx, const int32_t* __restrict__ cand_nvalid,
                         const int32_t* __restrict__ order, int kcand, K4Chunk ck,
                         const uint32_t* __restrict__ flags, uint32_t esc_limit, unsigned long long* __restrict__ dotacc) {
    static_assert(LPS >= 4 && LPS <= 32 && (LPS & (LPS - 1)) == 0, "<= 3 foreign elements per end go to lanes 0-2 of a group");
    if (!k4c_usable(flags, esc_limit)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t* tab = smem_raw;                                                            // [K4C_TAB]
    unsigned long long* dotc = reinterpret_cast<unsigned long long*>(smem_raw + K4C_TAB);   // [kcand] by sorted position
    uint32_t* sl = reinterpret_cast<uint32_t*>(dotc + kcand);                           // [kcand][sls] absolute positions
    const int sls = (wg + 1) | 1;                          // odd row stride: the 8 groups of a warp read 8 banks
    uint32_t* qdir = sl + (size_t)kcand * sls;                                          // [wg + 1]
    __shared__ int s_esc;

    const int64_t g = blockIdx.x / nq;
    const int64_t q = blockIdx.x - g * nq;
    const int64_t qrow = q_row0 + q;
    const int w0 = (int)g * wg;
    const int nwin = min(wg, n_windows - w0);
    const int64_t nw1 = (int64_t)n_windows + 1;
    const int lane = threadIdx.x & 31;
    constexpr int NG = T / LPS;                  // lane groups per CTA
    constexpr uint32_t ROUND = 4u * LPS * U;               // non-zeros per round of a lane group
    const int ncand = max(0, min((cand_idx ? cand_nvalid[q] : ck.n_all) - ck.cand_off, kcand));
    const int gi = threadIdx.x / LPS;
    const int ql = lane & (LPS - 1);
    const int gw = (threadIdx.x >> 5) * (32 / LPS);        // first lane group of this warp

    for (int w = threadIdx.x; w <= nwin; w += T) qdir[w] = __ldg(dir + qrow * nw1 + w0 + w);
    if (threadIdx.x == 0) s_esc = 0;
    __syncthreads();
    if (qdir[0] == qdir[nwin] || ncand == 0) return;       // the query row is empty in this whole group (uniform)

    const int64_t xb = indptr[qrow];
    auto next_window = [&](int j) { while (j < nwin && qdir[j + 1] == qdir[j]) ++j; return j; };
    uint32_t qn0 = 0u, qn1 = 0u;
    auto prefetch_keys = [&](int j) {
        if (j >= nwin) return;
        const uint32_t i0 = qdir[j] + threadIdx.x, e = qdir[j + 1];
        if (i0 < e) qn0 = __ldg(packed + xb + i0);
        if (i0 + T < e) qn1 = __ldg(packed + xb + i0 + T);
    };
    int j = next_window(0);
    prefetch_keys(j);

    for (int c = threadIdx.x; c < ncand; c += T) dotc[c] = 0ull;
    {
        const int per = nwin + 1, total = ncand * per;
        for (int x = threadIdx.x; x < total; x += T) {
            const int i = x / per, jj = x - i * per;
            const int c = __ldg(order + q * (int64_t)kcand + i);
            const int32_t d = cand_idx ? cand_idx[q * ck.cand_stride + ck.cand_off + c] : ck.cand_off + c;
            sl[i * sls + jj] = (uint32_t)(indptr[d] + (int64_t)__ldg(dir + (int64_t)d * nw1 + w0 + jj));
        }
    }
    for (int i = threadIdx.x; i < K4C_TAB / 16; i += T)
        reinterpret_cast<uint4*>(tab)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();                                       // sl complete

    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 v[U];                                            // one round of this lane
    uint32_t s_ys = 0u, s_ye = 0u, s_hw = 0u, s_tw = 0u;   // the slice v belongs to; its foreign head / tail word on this lane
    // one round: chunks bp[0], bp[LPS], ... of this lane while they start inside the slice (rem = non-zeros from this
    // lane's first chunk to the slice end; <= 0: nothing)
    auto load_round = [&](const uint4* __restrict__ bp, int rem) {
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = rem > 4 * LPS * u ? __ldg(bp + LPS * u) : zero4;
    };
    // slice of sorted candidate i in window jw: bounds, foreign words, round 0 (nothing is read for an empty slice)
    auto open_slice = [&](int i, int jw) {
        s_ys = 0u; s_ye = 0u; s_hw = 0u; s_tw = 0u;
        if (i < ncand && jw < nwin) {
            const uint32_t ys = sl[i * sls + jw], ye = sl[i * sls + jw + 1];
            if (ye > ys) { s_ys = ys; s_ye = ye; }
        }
        const uint32_t a0 = s_ys & ~3u;
        if (ql < (int)(s_ys - a0)) s_hw = __ldg(packed + a0 + ql);                 // in the first chunk, before the slice
        if (ql < (int)((0u - s_ye) & 3u)) s_tw = __ldg(packed + s_ye + ql);         // in the last chunk, behind the slice
        const uint32_t p0 = a0 + 4u * ql;
        load_round(reinterpret_cast<const uint4*>(packed + p0), (int)(s_ye - p0));
    };
    open_slice(gi, j);

    while (j < nwin) {
        const uint32_t k0 = qdir[j], k1 = qdir[j + 1];
        const uint32_t qc0 = qn0, qc1 = qn1;               // this window's keys (prefetched)
        __syncthreads();                                   // table clear / previous un-build done
        bool any_big = false;
        {
            const uint32_t i0 = k0 + threadIdx.x;
            if (i0 < k1) { const uint32_t x = qc0 & K4C_CNT_MASK; tab[qc0 >> K4C_CNT_BITS] = (uint8_t)min(x, 255u); any_big |= x >= 255u; }
            if (i0 + T < k1) { const uint32_t x = qc1 & K4C_CNT_MASK; tab[qc1 >> K4C_CNT_BITS] = (uint8_t)min(x, 255u); any_big |= x >= 255u; }
            for (uint32_t i = i0 + 2 * T; i < k1; i += T) {
                const uint32_t p = __ldg(packed + xb + i);
                const uint32_t x = p & K4C_CNT_MASK;
                tab[p >> K4C_CNT_BITS] = (uint8_t)min(x, 255u);
                any_big |= x >= 255u;
            }
        }
        if (any_big) s_esc = j + 1;                        // (j + 1 never repeats within a CTA: no reset needed)
        const int jn = next_window(j + 1);
        prefetch_keys(jn);                                 // consumed at the top of the next iteration
        __syncthreads();
        // ---- stream ---------------------------------------------------------------------------------------------
        {
            const bool esc = s_esc == j + 1;
            for (int ib = gw; ib < ncand; ib += NG) {      // warp-uniform trip count: the reduction below is warp-wide
                const int i = ib + (lane / LPS);
                if (ib != gw) open_slice(i, j);            // (only lists longer than NG candidates)
                unsigned long long acc = 0ull;
                const bool live = s_ye > s_ys;             // uniform in the lane group
                if (live) {
                    if (!esc) {
                        auto probe_round = [&](const uint4 (&x)[U]) {   // <= 4 U products < 2^24 each
                            uint32_t a32 = 0u;
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                a32 = k4c_mac(x[u].x, (uint32_t)tab[x[u].x >> K4C_CNT_BITS], a32);
                                a32 = k4c_mac(x[u].y, (uint32_t)tab[x[u].y >> K4C_CNT_BITS], a32);
                                a32 = k4c_mac(x[u].z, (uint32_t)tab[x[u].z >> K4C_CNT_BITS], a32);
                                a32 = k4c_mac(x[u].w, (uint32_t)tab[x[u].w >> K4C_CNT_BITS], a32);
                            }
                            return a32;
                        };
                        const uint32_t p0 = (s_ys & ~3u) + 4u * ql;
                        const uint4* __restrict__ bp = reinterpret_cast<const uint4*>(packed + p0);
                        int rem = (int)(s_ye - p0);
                        uint32_t left = s_ye - (s_ys & ~3u);           // non-zeros from the aligned start (group-uniform)
                        // Further rounds while the slice goes on. The loops are NOT unrolled: nvcc 12.9 unrolled the <8, 2>
                        // instance of an earlier form (loop condition on rem + 4 * ql) by four with the lane term's sign flipped
                        // in the unrolled exit test, and lanes 1.. of a group lost rounds.
                        // (a second register set that requests round r + 1 before round r is probed was measured: 10.7 ms
                        // against 10.4 ms -- the kernel is bound by L1 wavefronts, not by the latency of these loads)
                        acc = probe_round(v);
#pragma unroll 1
                        for (; left > ROUND; left -= ROUND) {
                            bp += LPS * U;
                            rem -= (int)ROUND;
                            load_round(bp, rem);
                            acc += probe_round(v);
                        }
                        uint32_t f32 = k4c_mac(s_hw, (uint32_t)tab[s_hw >> K4C_CNT_BITS], 0u);   // the foreign non-zeros again
                        f32 = k4c_mac(s_tw, (uint32_t)tab[s_tw >> K4C_CNT_BITS], f32);
                        acc -= f32;                        // (per lane this may wrap; the group's sum is exact mod 2^64)
                    } else {
                        // this window of the query holds counts >= 255: element-wise over the exact slice, table bytes of
                        // 255 looked up exactly (rare; v is not used)
                        const uint32_t* __restrict__ qp = packed + xb + k0;
                        const int qn = (int)(k1 - k0);
                        for (uint32_t p = s_ys + ql; p < s_ye; p += LPS) {
                            const uint32_t w = __ldg(packed + p);
                            const uint32_t t = w >> K4C_CNT_BITS, cnt = w & K4C_CNT_MASK;
                            const uint32_t tv = tab[t];
                            acc += (unsigned long long)(tv * cnt);
                            if (tv == 255u && cnt != 0u)
                                acc += (unsigned long long)(k4c_escape_value(qp, qn, t) - 255u) * (unsigned long long)cnt;
                        }
                    }
                }
#pragma unroll
                for (int o = 1; o < LPS; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (ql == 0 && live) dotc[i] += acc;       // only this lane group touches position i
            }
            open_slice(gi, jn);                            // in flight across un-build, the barriers and the next build
        }
        // ---- un-build: zero exactly the bytes this window set ----------------------------------------------------
        __syncthreads();
        {
            const uint32_t i0 = k0 + threadIdx.x;
            if (i0 < k1) tab[qc0 >> K4C_CNT_BITS] = 0;
            if (i0 + T < k1) tab[qc1 >> K4C_CNT_BITS] = 0;
            for (uint32_t i = i0 + 2 * T; i < k1; i += T) tab[__ldg(packed + xb + i) >> K4C_CNT_BITS] = 0;
        }
        j = jn;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncand; i += T)
        if (dotc[i] != 0ull) atomicAdd(&dotacc[q * (int64_t)kcand + __ldg(order + q * (int64_t)kcand + i)], dotc[i]);
}

// distances from the exact integer dot products, (distance asc, index asc) order, keep k
__global__ void __launch_bounds__(256)
k4c_finalize_kernel(const uint32_t* __restrict__ flags, uint32_t esc_limit, const unsigned long long* __restrict__ dotacc,
                    const double* __restrict__ norm2, int64_t q_row0, const int32_t* __restrict__ cand_idx,
                    const int32_t* __restrict__ cand_nvalid, int kcand, int P, int k, K4Chunk ck,
                    int32_t* __restrict__ idx_out, float* __restrict__ dist_out, int32_t* __restrict__ nvalid_out) {
    if (!k4c_usable(flags, esc_limit)) return;
    extern __shared__ unsigned long long fkeys[];
    const int64_t q = blockIdx.x;
    const int64_t qrow = q_row0 + q;
    const int ncand = max(0, min((cand_idx ? cand_nvalid[q] : ck.n_all) - ck.cand_off, kcand));
    const double nq2 = norm2[qrow];
    for (int c = threadIdx.x; c < P; c += blockDim.x) {
        unsigned long long key = ~0ull;
        if (c < ncand) {
            const int32_t d = cand_idx ? cand_idx[q * ck.cand_stride + ck.cand_off + c] : ck.cand_off + c;
            // integers below 2^53: every term and the result are exact in fp64
            double d2 = d == qrow ? 0.0 : nq2 + norm2[d] - 2.0 * (double)dotacc[q * (int64_t)kcand + c];
            if (d2 < 0.0) d2 = 0.0;
            const float dist = (float)sqrt(d2);
            key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(uint32_t)d;
        }
        fkeys[c] = key;
    }
    bitonic_sort_shared(fkeys, P);
    const int nv = min(ncand, k);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        int32_t oi = -1;
        float od = __int_as_float(0x7f800000);
        if (t < nv) {
            const unsigned long long key = fkeys[t];
            oi = (int32_t)(uint32_t)(key & 0xFFFFFFFFull);
            od = __uint_as_float((uint32_t)(key >> 32));
        }
        idx_out[q * ck.out_stride + ck.out_off + t] = oi;
        dist_out[q * ck.out_stride + ck.out_off + t] = od;
    }
    if (threadIdx.x == 0 && ck.write_nvalid) nvalid_out[q] = nv;
}

// windows per group: the group's share of the packed stream (4 B x nnz x wg / n_windows) should fit the part of L2 we
// want it to live in (group_bytes), and the per-candidate window offsets must fit in shared memory
int k4c_windows_per_group(int n_windows, int64_t db_nnz, int kcand, int64_t group_bytes) {
    const int64_t smem_left = (int64_t)(225 * 1024 / K4C_CTAS - 1024) - K4C_TAB - (int64_t)kcand * 8 - 64;
    if (smem_left < (int64_t)kcand * 8) return -1;
    int64_t wg_smem = smem_left / ((int64_t)kcand * 4) - 2;                  // (wg + 2) * 4 * kcand + (wg + 1) * 4 <= left
    while (wg_smem > 1 && (wg_smem + 2) * 4 * (int64_t)kcand + (wg_smem + 1) * 4 > smem_left) --wg_smem;
    if (wg_smem < 1) return -1;
    int64_t wg = n_windows;
    const int64_t stream_bytes = 4 * std::max<int64_t>(db_nnz, 1);
    if (group_bytes < 0) wg = -group_bytes;                                  // forced (tests, experiments)
    else if (group_bytes > 0 && stream_bytes > group_bytes) wg = std::max<int64_t>(1, group_bytes * n_windows / stream_bytes);
    wg = std::min<int64_t>(std::min<int64_t>(wg, wg_smem), n_windows);
    return (int)wg;
}

// scratch of one launch (in uint64 units): nq * kcand dot products, nq * kcand int32 of candidate order, nq * K4C_THREADS
// uint16 of lane map
size_t k4c_dotacc_elems(int64_t nq, int kcand) {
    return (size_t)nq * (size_t)kcand + ((size_t)nq * (size_t)kcand + 1) / 2 + ((size_t)nq * K4C_THREADS + 3) / 4 + 2;
}

// which of the two counts kernels launch_rerank_counts uses (SCHIC_K4_KERNEL = direct | iter; lane-group experiments
// SCHIC_K4_LPS = 0 | 2 exist in the iterator kernel only)
bool k4c_direct_selected() {
    const char* kk = getenv("SCHIC_K4_KERNEL");
    if (kk && strcmp(kk, "iter") == 0) return false;
    if (const char* e = getenv("SCHIC_K4_LPS")) { const int v = atoi(e); if (v != 4 && v != 8) return false; }
    return true;
}
const char* k4c_kernel_name() { return k4c_direct_selected() ? "k4c_rerank_direct_kernel" : "k4c_rerank_kernel"; }

// One launch (pair) over candidate columns [cand_off, cand_off + kcand); same output contract as launch_rerank.
// d_dotacc: >= nq * kcand uint64 scratch (zeroed here). Returns launches, or -1 when kcand does not fit.
int launch_rerank_counts(const int64_t* d_indptr, const uint32_t* d_packed, const double* d_norm2, const uint32_t* d_dir,
                         int n_windows, int64_t db_nnz, int64_t group_bytes, int64_t q_row0, int64_t nq,
                         const int32_t* d_cand_idx, const int32_t* d_cand_nvalid, int cand_stride, int cand_off, int kcand,
                         int k, int32_t* d_idx_out, float* d_dist_out, int out_stride, int out_off, int final,
                         int32_t* d_nvalid_out, int n_all, const uint32_t* d_flags, uint32_t esc_limit,
                         unsigned long long* d_dotacc, cudaStream_t stream) {
    if (nq <= 0) return 0;
    if (kcand > 1024 || db_nnz >= 0xFFFFFFF0ll) return -1;      // (stream positions are kept as 32-bit in shared memory)
    const int wg = k4c_windows_per_group(n_windows, db_nnz, kcand, group_bytes);
    if (wg < 1) return -1;
    const int64_t groups = (n_windows + wg - 1) / wg;
    if (groups * nq > 0x7FFFFFFFll) return -1;
    const K4Chunk ck{cand_stride, cand_off, out_stride, out_off, final, n_all};
    // (the direct kernel pads the rows of the slice table to an odd stride: wg + 2 words at most)
    const size_t smem = (size_t)K4C_TAB + (size_t)kcand * 8 + (size_t)kcand * (wg + 2) * 4 + (size_t)(wg + 1) * 4 + 16;
    if (smem > (size_t)(225 * 1024 / K4C_CTAS - 1024)) return -1;
    if (cudaMemsetAsync(d_dotacc, 0, (size_t)nq * kcand * sizeof(unsigned long long), stream) != cudaSuccess) return -1;
    int32_t* d_order = reinterpret_cast<int32_t*>(d_dotacc + (size_t)nq * kcand);
    uint16_t* d_lane_map = reinterpret_cast<uint16_t*>(d_order + (size_t)nq * kcand);
    const int P = next_pow2(kcand);
    // lanes per slice: fixed groups of 4 (8 for short lists) dealt round-robin in length-sorted order. SCHIC_K4_LPS = 4 | 8
    // forces one; 0 = group sizes proportional to the candidates' row lengths (lists of <= 128; measured at S10:
    // 15.7 ms against 15.1 ms for groups of 4 and 16.5 ms for groups of 8 -- the barrier structure, not the slice
    // imbalance, is what keeps issue utilisation at ~52 %; kept as an experiment switch)
    int lps = kcand > K4C_THREADS / 8 ? 4 : 8;
    if (const char* e = getenv("SCHIC_K4_LPS")) { const int v = atoi(e); lps = v == 8 ? 8 : (v == 4 ? 4 : (v == 2 ? 2 : (kcand <= K4C_THREADS / 4 ? 0 : 4))); }
    k4c_order_kernel<<<(unsigned)nq, 256, (size_t)P * 8, stream>>>(d_indptr, d_cand_idx, d_cand_nvalid, kcand, P, ck, d_flags,
                                                                   esc_limit, d_order, lps == 0 ? d_lane_map : nullptr);
    // SCHIC_K4_KERNEL = direct (default) | iter (the batch-iterator kernel); SCHIC_K4_U = chunks per lane and round
    const bool direct = k4c_direct_selected() && lps >= 4;
    if (direct) {
        // chunks per lane and round: 4 x 4 lanes x 4 = 64 non-zeros cover most S10 slices in one or two rounds (measured
        // 3 / 4 / 5 / 6 / 8 chunks: 11.6 / 10.4 / 10.4 / 10.6 / 13.0 ms); threads: the smallest build that gives every
        // candidate its lane group
        int u = lps == 8 ? 3 : 4;
        if (const char* e = getenv("SCHIC_K4_U")) u = atoi(e);
        int threads = kcand * lps <= 416 ? 416 : 512;
        if (const char* e = getenv("SCHIC_K4_THREADS")) threads = atoi(e) == 416 ? 416 : 512;
        void (*kern)(const int64_t*, const uint32_t*, const uint32_t*, int, int, int64_t, int64_t, const int32_t*, const int32_t*,
                     const int32_t*, int, K4Chunk, const uint32_t*, uint32_t, unsigned long long*);
        if (threads == 416)
            kern = lps == 8 ? (u <= 2 ? k4c_rerank_direct_kernel<8, 2, 416> : k4c_rerank_direct_kernel<8, 3, 416>)
                            : (u <= 4 ? k4c_rerank_direct_kernel<4, 4, 416> : k4c_rerank_direct_kernel<4, 6, 416>);
        else
            kern = lps == 8 ? (u <= 2 ? k4c_rerank_direct_kernel<8, 2, 512> : k4c_rerank_direct_kernel<8, 3, 512>)
                            : (u <= 4 ? k4c_rerank_direct_kernel<4, 4, 512> : k4c_rerank_direct_kernel<4, 6, 512>);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<(unsigned)(groups * nq), threads, smem, stream>>>(
            d_indptr, d_packed, d_dir, n_windows, wg, q_row0, nq, d_cand_idx, d_cand_nvalid, d_order, kcand, ck,
            d_flags, esc_limit, d_dotacc);
    } else {
        auto kern = lps == 0 ? k4c_rerank_kernel<0> : (lps == 2 ? k4c_rerank_kernel<2> : (lps == 4 ? k4c_rerank_kernel<4> : k4c_rerank_kernel<8>));
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<(unsigned)(groups * nq), K4C_THREADS, smem, stream>>>(
            d_indptr, d_packed, d_dir, n_windows, wg, q_row0, nq, d_cand_idx, d_cand_nvalid, d_order, d_lane_map, kcand, ck,
            d_flags, esc_limit, d_dotacc);
    }
    k4c_finalize_kernel<<<(unsigned)nq, 256, (size_t)P * 8, stream>>>(d_flags, esc_limit, d_dotacc, d_norm2, q_row0,
                                                                      d_cand_idx, d_cand_nvalid, kcand, P, k, ck, d_idx_out,
                                                                      d_dist_out, d_nvalid_out);
    return 3;
}

}  // namespace schic
