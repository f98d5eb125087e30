// k1_signatures.cu -- K1: MinHash signatures, sm_100a.
//
// Replaces the signature loop of sparse_neighbors_search.MinHash.fit / partial_fit
// (call sites /root/reference/schicexplorer/scHicClusterMinHash.py:351, :353, :406; algorithm in
// docs/content/Example_analysis.rst:466 and docs/content/tools/scHicClusterMinHash.rst:11):
//     sig[c][j] = min over non-zero features f of cell c of  wang32((f+1)*(j+1)) mod (2^31-1)
//
// Mapping (B200-first, not the reference's loop nest):
//   * The CSR index array is cut into fixed chunks of K1_CHUNK features, one CTA per chunk, so
//     every CTA does the same amount of work whatever the row-length skew (1.6k..60k nnz/cell).
//     A chunk may straddle rows; each (row, sub-range) segment is reduced separately.
//   * The CTA stages its chunk once into shared memory with coalesced (128-bit when aligned)
//     loads, already multiplied: g = (f+1)*32767. Threads then read it as an LDS.128 broadcast.
//   * lane <-> hash function. A warp owns R*32 hash functions ("R slabs"); each lane keeps its R
//     running minima in registers for the whole segment, so the min-reduction needs no shuffles
//     and no shared memory -- one feature load feeds 32*R hash evaluations.
//   * Integer strength reduction (all exact in uint32 wraparound arithmetic):
//       - ~k + (k<<15) == 32767*k - 1, and k = (f+1)(j+1), so the first mix step is
//         a = g*(j+1) - 1 = g*(32 r) + [g*(j0+1) - 1]: one IMAD with an immediate per evaluation.
//       - "x mod (2^31-1), then min" is replaced by two running minima of k' = k+2, one unsigned
//         and one signed. The three residue segments of k' are [0,2) (k = 2M, 2M+1 -> 0, 1),
//         [2, M+2) (value k'-2) and [M+2, 2^32) (value k'-M-2); umin finds the minimum of the
//         first non-empty segment, smin the minimum of the last one. The only ambiguous outcomes
//         (umin < 2, or smin == INT_MIN, probability 3/2^32 per evaluation) take an exact slow
//         path over the segment, so the result is bit-identical to mod-then-min.
//     Per evaluation: 4 ops on the FMA pipe (IMAD) + 8 on the ALU pipe (3x SHF+LOP3, 2x MNMX).
//   * Partial minima of a segment are merged with atomicMin (the signature rows are pre-filled
//     with the sentinel M, which is also the result for empty rows).
#include "common.cuh"

namespace schic {

constexpr int K1_CHUNK = 2048;    // features per CTA (8 KB of shared memory)
constexpr int K1_MAX_WARPS = 8;

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, int64_t n, uint32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

__device__ __forceinline__ uint32_t mix_from_a(uint32_t a) {
    // a = ~k + (k << 15) already applied
    uint32_t b = a ^ (a >> 12);
    uint32_t c = b * 5u;
    uint32_t d = c ^ (c >> 4);
    uint32_t e = d * 2057u;
    return e ^ (e >> 16);
}

// Exact (mod, then min) over a segment for one hash function; only used on the rare ambiguous
// outcomes of the two-minima trick.
__device__ __noinline__ uint32_t exact_segment_min(const uint32_t* __restrict__ sg, int len, uint32_t jp1) {
    uint32_t best = kModulus;
    for (int i = 0; i < len; ++i) {
        const uint32_t k = mix_from_a(sg[i] * jp1 - 1u);
        const uint32_t v = k % kModulus;
        best = v < best ? v : best;
    }
    return best;
}

// Instruction-selection variant V (bit field), chosen from measurements on the B200:
//   bit 0     the "+2" is issued as IMAD (FMA pipe) instead of VIADD (ALU pipe)
//   bit 1     features are consumed in pairs so both running minima use the 3-input VIMNMX3
//   bits 2-3  number (0..3) of the three ">> s" done as IMAD.HI (FMA pipe) instead of SHF (ALU pipe)
//   bit 4     the final ">> 16" comes from IMAD.WIDE (FMA pipe) + a masking LOP3
template <int V>
__device__ __forceinline__ uint32_t shr_v(uint32_t x, int s, int which) {
    if (which < ((V >> 2) & 3)) {
        uint32_t r;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(1u << (32 - s)));
        return r;
    }
    return x >> s;
}

template <int V>
__device__ __forceinline__ uint32_t kprime(uint32_t a, uint32_t one) {
    // a = ~k + (k << 15) already applied; returns wang32 output + 2
    const uint32_t b = a ^ shr_v<V>(a, 12, 2);
    const uint32_t c = b * 5u;
    const uint32_t d = c ^ shr_v<V>(c, 4, 1);
    const uint32_t e = d * 2057u;
    uint32_t k;
    if (V & 16) {
        // (e >> 16) from the FMA pipe: the high word of d * (2057 << 16) holds bits 16..47 of d*2057,
        // whose low 16 bits are e >> 16; one LOP3 masks and xors.
        const unsigned long long wide = (unsigned long long)d * (unsigned long long)(2057u << 16);
        k = e ^ ((uint32_t)(wide >> 32) & 0xFFFFu);
    } else {
        k = e ^ shr_v<V>(e, 16, 0);
    }
    if (V & 1) {
        uint32_t kp;
        asm("mad.lo.u32 %0, %1, %2, 2;" : "=r"(kp) : "r"(k), "r"(one));
        return kp;
    }
    return k + 2u;
}

template <int R, int V>
__device__ __forceinline__ void segment_unit(const uint32_t* __restrict__ sg, int len, int j0 /* lane's first hash */,
                                             int H, uint32_t one, uint32_t* __restrict__ sig_row) {
    uint32_t mU[R];
    int32_t mS[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { mU[r] = 0xFFFFFFFFu; mS[r] = 0x7FFFFFFF; }
    const uint32_t jp1 = (uint32_t)j0 + 1u;

    auto eval1 = [&](uint32_t g) {
        const uint32_t base = g * jp1 - 1u;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t kp = kprime<V>(g * (uint32_t)(32 * r) + base, one);
            mU[r] = min(mU[r], kp);
            mS[r] = min(mS[r], (int32_t)kp);
        }
    };
    auto eval2 = [&](uint32_t ga, uint32_t gb) {
        if (V & 2) {
            const uint32_t base_a = ga * jp1 - 1u, base_b = gb * jp1 - 1u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t ka = kprime<V>(ga * (uint32_t)(32 * r) + base_a, one);
                const uint32_t kb = kprime<V>(gb * (uint32_t)(32 * r) + base_b, one);
                mU[r] = min(mU[r], min(ka, kb));
                mS[r] = min(mS[r], min((int32_t)ka, (int32_t)kb));
            }
        } else {
            eval1(ga); eval1(gb);
        }
    };

    // sg is 16-byte aligned only at multiples of 4 from the chunk start: peel to alignment.
    int i = 0;
    const int mis = (int)(((uintptr_t)sg >> 2) & 3);
    const int head = min(len, (4 - mis) & 3);
    for (; i < head; ++i) eval1(sg[i]);
    for (; i + 4 <= len; i += 4) {
        const uint4 g4 = *reinterpret_cast<const uint4*>(sg + i);
        eval2(g4.x, g4.y); eval2(g4.z, g4.w);
    }
    for (; i < len; ++i) eval1(sg[i]);

#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = j0 + 32 * r;
        if (j < H) {
            uint32_t v;
            if (mU[r] < 2u || mS[r] == (int32_t)0x80000000) {
                v = exact_segment_min(sg, len, (uint32_t)j + 1u);
            } else {
                const uint32_t valA = (mU[r] <= kModulus + 1u) ? mU[r] - 2u : mU[r] - (kModulus + 2u);
                const uint32_t valB = (mS[r] < 0) ? (uint32_t)mS[r] - (kModulus + 2u) : kModulus;
                v = min(valA, valB);
            }
            atomicMin(sig_row + j, v);
        }
    }
}

// grid = ceil(nnz / K1_CHUNK); block = 32 * n_warps; `units` = ceil(ceil(H/32) / R).
// `one` is the constant 1 passed at run time so that ptxas keeps the IMAD of variant bit 0.
template <int R, int V>
__global__ void __launch_bounds__(32 * K1_MAX_WARPS)
k1_signatures32_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat, int64_t n_rows,
                       int64_t nnz, int64_t indptr0, int H, int units, uint32_t one, uint32_t* __restrict__ sig) {
    __shared__ __align__(16) uint32_t sg[K1_CHUNK];
    __shared__ int64_t s_first_row;

    const int64_t p0 = (int64_t)blockIdx.x * K1_CHUNK;           // chunk = [p0, p1) relative to indptr0
    const int64_t p1 = min(p0 + (int64_t)K1_CHUNK, nnz);
    const int len = (int)(p1 - p0);
    const uint32_t* __restrict__ src = feat + indptr0 + p0;

    // stage: g = (f+1)*32767, coalesced; 128-bit loads when the source is 16-byte aligned
    if ((((uintptr_t)src) & 15) == 0) {
        const uint4* __restrict__ src4 = reinterpret_cast<const uint4*>(src);
        for (int i = threadIdx.x; i < (len >> 2); i += blockDim.x) {
            uint4 f = __ldg(src4 + i);
            f.x = (f.x + 1u) * 32767u; f.y = (f.y + 1u) * 32767u;
            f.z = (f.z + 1u) * 32767u; f.w = (f.w + 1u) * 32767u;
            reinterpret_cast<uint4*>(sg)[i] = f;
        }
        for (int i = (len & ~3) + threadIdx.x; i < len; i += blockDim.x) sg[i] = (__ldg(src + i) + 1u) * 32767u;
    } else {
        for (int i = threadIdx.x; i < len; i += blockDim.x) sg[i] = (__ldg(src + i) + 1u) * 32767u;
    }
    // first row whose end lies beyond p0: smallest r with indptr[r+1] - indptr0 > p0
    if (threadIdx.x == 0) {
        int64_t lo = 0, hi = n_rows - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (indptr[mid + 1] - indptr0 > p0) hi = mid; else lo = mid + 1;
        }
        s_first_row = lo;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    int64_t row = s_first_row;
    int64_t seg0 = p0;
    while (seg0 < p1) {
        int64_t row_end = indptr[row + 1] - indptr0;
        while (row_end <= seg0) { ++row; row_end = indptr[row + 1] - indptr0; }   // skip empty rows
        const int64_t seg1 = min(row_end, p1);
        const uint32_t* s = sg + (int)(seg0 - p0);
        const int slen = (int)(seg1 - seg0);
        uint32_t* sig_row = sig + row * (int64_t)H;
        for (int u = warp; u < units; u += n_warps)
            segment_unit<R, V>(s, slen, u * R * 32 + lane, H, one, sig_row);
        seg0 = seg1;
    }
}

// ---- filtered variant ---------------------------------------------------------------------------
// The kernel above evaluates the full hash + both running minima for every (feature, hash function):
// 7 ALU-pipe ops per evaluation, and ncu shows the ALU pipe 98 % busy. But after m features of a row
// the probability that the next one lowers a given minimum is only ~1/m, so almost every evaluation
// only has to prove "not smaller than the current minimum". The last mix step is
//     k6 = k5 ^ (k5 >> 16),   v = k6 mod (2^31 - 1)   with   k5 = d * 2057,
// and bits 30..16 of k6 equal bits 30..16 of k5: a 15-bit prefix of v (up to the +1 carry of the
// modulo fold) is known one xor-shift and the whole modulo earlier. With
//     t = d * 4114 + 0x20000  (mod 2^32)          -- one IMAD: 2*k5, prefix moved up by one step
// v < best implies t <= thr(best) = ((best >> 16) + 2) * 2^17 - 1 (the +0x20000 makes the three
// residue-wrapping inputs, prefix 0x7FFF, land on t < 2^17 so they always pass; proof in DESIGN.md).
// Fast path per evaluation: 3 IMAD (FMA pipe) + 2 x (SHF + LOP3) + 1 compare = 5 ALU-pipe ops instead
// of 7; only when some lane passes the filter (true new minima: ~32 ln(m) per warp and hash register
// over a row, plus 2^-14 false positives) the warp takes an exact slow path for that pair of features.
// To make the minima long-lived a CTA owns a whole row (or an up-to-K1F_PART slice of a long row),
// streaming it through a shared-memory tile; work items are sorted by length (longest first) by a
// tiny single-CTA kernel so the tail of the grid is made of short items.
constexpr int K1F_TILE = 2048;
constexpr int K1F_PART = 16384;
constexpr int K1F_BINS = 64;

struct K1Item { int64_t begin; int32_t len; int32_t row; };

// items[]: every row cut into ceil(len / K1F_PART) equal parts, ordered by part length, longest first
// (counting sort on 64 length bins; order inside a bin is arbitrary, the result does not depend on it).
__global__ void __launch_bounds__(1024)
k1f_build_items_kernel(const int64_t* __restrict__ indptr, int64_t n_rows, K1Item* __restrict__ items,
                       int* __restrict__ n_items) {
    __shared__ int hist[K1F_BINS];
    __shared__ int cursor[K1F_BINS];
    for (int i = threadIdx.x; i < K1F_BINS; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    auto split = [](int64_t len, int& nparts, int& plen, int& bin) {
        nparts = (int)((len + K1F_PART - 1) / K1F_PART);
        plen = nparts ? (int)((len + nparts - 1) / nparts) : 0;
        bin = K1F_BINS - 1 - (int)(((int64_t)plen * (K1F_BINS - 1)) / K1F_PART);    // long parts -> low bins
    };
    for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
        int np, pl, b;
        split(indptr[r + 1] - indptr[r], np, pl, b);
        if (np) atomicAdd(&hist[b], np);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < K1F_BINS; ++b) { cursor[b] = acc; acc += hist[b]; }
        *n_items = acc;
    }
    __syncthreads();
    for (int64_t r = threadIdx.x; r < n_rows; r += blockDim.x) {
        int np, pl, b;
        const int64_t beg = indptr[r], len = indptr[r + 1] - beg;
        split(len, np, pl, b);
        if (!np) continue;
        const int pos = atomicAdd(&cursor[b], np);
        for (int p = 0; p < np; ++p) {
            const int64_t o = (int64_t)p * pl;
            K1Item it;
            it.begin = beg + o;
            it.len = (int)min((int64_t)pl, len - o);
            it.row = (int)r;
            items[pos + p] = it;
        }
    }
}

__device__ __forceinline__ uint32_t k1f_filter_key(uint32_t a) {
    const uint32_t b = a ^ (a >> 12);
    const uint32_t c = b * 5u;
    const uint32_t d = c ^ (c >> 4);
    return d * 4114u + 0x20000u;
}

__device__ __forceinline__ uint32_t k1f_threshold(uint32_t best) {
    const uint32_t hb = best >> 16;
    return hb >= 0x7FFEu ? 0xFFFFFFFFu : ((hb + 2u) << 17) - 1u;
}

__device__ __forceinline__ void k1f_exact_update(uint32_t a, uint32_t& best, uint32_t& thr) {
    const uint32_t k6 = mix_from_a(a);
    uint32_t v = (k6 & kModulus) + (k6 >> 31);     // k6 mod (2^31 - 1): fold the top bit, then one subtract
    v = v >= kModulus ? v - kModulus : v;
    if (v < best) { best = v; thr = k1f_threshold(v); }
}

// Exact (mod, then min) over a whole item for one hash constant, features read straight from global
// memory (all lanes the same address): only used when the speculative start threshold was too low
// for some lane (probability ~e^-11 per lane and hash function).
__device__ __noinline__ uint32_t k1f_exact_item_min(const uint32_t* __restrict__ src, int len, uint32_t c) {
    uint32_t best = kModulus;
    for (int i = 0; i < len; ++i) {
        const uint32_t k6 = mix_from_a((__ldg(src + i) + 1u) * c - 1u);
        uint32_t v = (k6 & kModulus) + (k6 >> 31);
        v = v >= kModulus ? v - kModulus : v;
        best = v < best ? v : best;
    }
    return best;
}

// grid = (item upper bound, unit groups); block = 32 * warps, one unit (R*32 hash functions) per warp.
// Per lane and hash register r: best[r] (exact running minimum), thr[r] (filter threshold), mn[r]
// (running minimum of the filter keys since the last exact step). The inner loop only feeds mn[r]
// (one 3-input min per two evaluations); after every 4 features one compare per r decides whether
// the exact path has to look at those 4 features.
// The start threshold is speculative: thr0 ~ 2^32 * 12 / len, i.e. the filter assumes the final
// minimum is below the level ~12 of the len features reach. The assumption is verified at the end
// (thr(best) <= thr0 proves that nothing that was skipped could have been smaller); where it fails
// the lane's minimum is recomputed exactly. This removes the ~ln(len) early updates per minimum.
// One group of 4 features for all R hash registers of the lane: feed the filter minima, then (rarely)
// the exact path. fp1[] = feature id + 1.
template <int R>
__device__ __forceinline__ void k1f_step4(const uint32_t (&fp1)[4], const uint32_t (&c)[R], uint32_t (&best)[R],
                                          uint32_t (&thr)[R], uint32_t (&mn)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t t0k = k1f_filter_key(fp1[0] * c[r] - 1u), t1k = k1f_filter_key(fp1[1] * c[r] - 1u);
        const uint32_t t2k = k1f_filter_key(fp1[2] * c[r] - 1u), t3k = k1f_filter_key(fp1[3] * c[r] - 1u);
        mn[r] = min(mn[r], min(t0k, t1k));
        mn[r] = min(mn[r], min(t2k, t3k));
    }
    bool any = false;
#pragma unroll
    for (int r = 0; r < R; ++r) any = any || (mn[r] <= thr[r]);
    if (any) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (mn[r] <= thr[r]) {
#pragma unroll
                for (int u = 0; u < 4; ++u) k1f_exact_update(fp1[u] * c[r] - 1u, best[r], thr[r]);
                mn[r] = 0xFFFFFFFFu;
            }
        }
    }
}

// grid = (item upper bound, unit groups); block = 32 * warps, one unit (R*32 hash functions) per warp.
// Per lane and hash register r: best[r] (exact running minimum), thr[r] (filter threshold), mn[r]
// (running minimum of the filter keys since the last exact step). The inner loop only feeds mn[r]
// (one 3-input min per two evaluations); after every 4 features one compare per r decides whether
// the exact path has to look at those 4 features.
// The start threshold is speculative: thr0 ~ 2^32 * 12 / len, i.e. the filter assumes the final
// minimum is below the level ~12 of the len features reach. The assumption is verified at the end
// (thr(best) <= thr0 proves that nothing that was skipped could have been smaller); where it fails
// the lane's minimum is recomputed exactly. This removes the ~ln(len) early updates per minimum.
// DIRECT = false: the CTA stages 2048-feature tiles in shared memory (two barriers per tile).
// DIRECT = true : no shared memory and no barriers -- every warp reads the features itself with
//                 warp-uniform loads (one L1 line serves 32 features of all warps of the CTA), one
//                 group of 4 prefetched into registers while the previous one is hashed.
template <int R, bool DIRECT>
__global__ void __launch_bounds__(32 * K1_MAX_WARPS)
k1_filter_kernel(const K1Item* __restrict__ items, const int* __restrict__ n_items, const uint32_t* __restrict__ feat,
                 int H, int units, uint32_t one /* = 1 at run time: keeps the "+1" an IMAD on the FMA pipe */,
                 uint32_t* __restrict__ sig) {
    __shared__ __align__(16) uint32_t tile[DIRECT ? 4 : K1F_TILE];
    if ((int)blockIdx.x >= *n_items) return;
    const K1Item it = items[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const int unit = blockIdx.y * n_warps + warp;
    const bool active = unit < units;
    const int j0 = unit * R * 32 + lane;

    const uint32_t thr0 = it.len <= 12 ? 0xFFFFFFFFu
                                       : (uint32_t)min((unsigned long long)0xFFFFFFFFu, (12ull << 32) / (unsigned)it.len);
    uint32_t c[R], best[R], thr[R], mn[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        c[r] = 32767u * (uint32_t)(j0 + 32 * r + 1);     // a = (f+1)(j+1)*32767 - 1 = (f+1)*c - 1
        best[r] = kModulus;
        thr[r] = thr0;
        mn[r] = 0xFFFFFFFFu;
    }
    const uint32_t* __restrict__ src = feat + it.begin;
    if (DIRECT) {
        if (!active) return;
        const int n4 = it.len & ~3;
        uint32_t nxt[4] = {0u, 0u, 0u, 0u};
        if (n4 > 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u) nxt[u] = __ldg(src + u);
        }
        for (int i = 0; i < n4; i += 4) {
            uint32_t cur[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) asm("mad.lo.u32 %0, %1, %2, 1;" : "=r"(cur[u]) : "r"(nxt[u]), "r"(one));
            if (i + 8 <= n4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) nxt[u] = __ldg(src + i + 4 + u);
            }
            k1f_step4<R>(cur, c, best, thr, mn);
        }
        if (n4 < it.len) {                                // 1..3 left: pad with copies of the last feature
            uint32_t cur[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) cur[u] = __ldg(src + min(n4 + u, it.len - 1)) + 1u;
            k1f_step4<R>(cur, c, best, thr, mn);
        }
    } else {
        for (int t0 = 0; t0 < it.len; t0 += K1F_TILE) {
            const int tlen = min(K1F_TILE, it.len - t0);
            const int tlen4 = (tlen + 3) & ~3;           // padded with copies of the last feature (min is idempotent)
            __syncthreads();
            for (int i = threadIdx.x; i < tlen4; i += blockDim.x) tile[i] = __ldg(src + t0 + min(i, tlen - 1)) + 1u;
            __syncthreads();
            if (!active) continue;
            for (int i = 0; i < tlen4; i += 4) {
                const uint4 f4 = *reinterpret_cast<const uint4*>(tile + i);
                const uint32_t cur[4] = {f4.x, f4.y, f4.z, f4.w};
                k1f_step4<R>(cur, c, best, thr, mn);
            }
        }
    }
    if (active) {
        uint32_t* __restrict__ sig_row = sig + (int64_t)it.row * H;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int j = j0 + 32 * r;
            // the speculation holds iff everything skipped under thr0 is provably >= best
            if (k1f_threshold(best[r]) > thr0) best[r] = k1f_exact_item_min(src, it.len, c[r]);
            if (j < H) atomicMin(sig_row + j, best[r]);
        }
    }
}

int64_t k1f_item_capacity(int64_t n_rows, int64_t nnz) { return n_rows + nnz / K1F_PART + 1; }

// scratch: K1Item[k1f_item_capacity] followed by one int (16-byte aligned)
size_t k1f_scratch_bytes(int64_t n_rows, int64_t nnz) {
    return (size_t)k1f_item_capacity(n_rows, nnz) * sizeof(K1Item) + 16;
}

static int pick_R_filter(int slabs) {
    int best = 5, best_waste = 1 << 30;
    for (int R = 6; R >= 4; --R) {
        const int waste = ((slabs + R - 1) / R) * R - slabs;
        if (waste < best_waste) { best_waste = waste; best = R; }
    }
    return best;
}

int launch_signatures32_filter(const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows, int64_t nnz,
                               int H, uint32_t* d_sig, void* d_scratch, int direct, cudaStream_t stream) {
    if (n_rows <= 0) return 0;
    const int64_t total = n_rows * (int64_t)H;
    const int64_t want = (total + 1023) / 1024;
    fill_u32_kernel<<<(int)(want < 148 * 8 ? want : 148 * 8), 256, 0, stream>>>(d_sig, total, kModulus);
    if (nnz <= 0) return 1;
    const int64_t cap = k1f_item_capacity(n_rows, nnz);
    K1Item* items = reinterpret_cast<K1Item*>(d_scratch);
    int* n_items = reinterpret_cast<int*>(items + cap);
    k1f_build_items_kernel<<<1, 1024, 0, stream>>>(d_indptr, n_rows, items, n_items);
    const int slabs = (H + 31) / 32;
    const int R = pick_R_filter(slabs);
    const int units = (slabs + R - 1) / R;
    int n_warps = units < K1_MAX_WARPS ? units : K1_MAX_WARPS;
    for (int w = K1_MAX_WARPS; w >= 4; --w)
        if (units % w == 0) { n_warps = w; break; }
    const dim3 grid((unsigned)cap, (unsigned)((units + n_warps - 1) / n_warps));
    const dim3 block(32 * n_warps);
#define K1F_LAUNCH(RR)                                                                                       \
    do {                                                                                                     \
        if (direct) k1_filter_kernel<RR, true><<<grid, block, 0, stream>>>(items, n_items, d_feat, H, units, 1u, d_sig);  \
        else k1_filter_kernel<RR, false><<<grid, block, 0, stream>>>(items, n_items, d_feat, H, units, 1u, d_sig);        \
    } while (0)
    switch (R) {
        case 4: K1F_LAUNCH(4); break;
        case 5: K1F_LAUNCH(5); break;
        default: K1F_LAUNCH(6); break;
    }
#undef K1F_LAUNCH
    return 3;
}

// ---- width 64 (alternative HashSpec): the same formula in uint64 before the modulo --------
__device__ __forceinline__ uint32_t hash64_dev(uint64_t f1, uint32_t jp1) {
    uint64_t k = f1 * (uint64_t)jp1;
    k = ~k + (k << 15);
    k ^= k >> 12;
    k += k << 2;
    k ^= k >> 4;
    k *= 2057ull;
    k ^= k >> 16;
    // k mod (2^31-1) by folding 31-bit limbs (2^31 == 1 mod M)
    uint64_t s = (k & kModulus) + ((k >> 31) & kModulus) + (k >> 62);
    s = (s & kModulus) + (s >> 31);
    uint32_t v = (uint32_t)s;
    return v >= kModulus ? v - kModulus : v;
}

// One CTA per (row, slab group): lane <-> hash, loop over the row's features (broadcast loads).
__global__ void __launch_bounds__(256)
k1_signatures64_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat_lo,
                       const uint32_t* __restrict__ feat_hi, int H,
                       uint32_t* __restrict__ sig) {
    const int64_t row = blockIdx.x;
    const int64_t b = indptr[row], e = indptr[row + 1];
    for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < H; j += gridDim.y * blockDim.x) {
        uint32_t best = kModulus;
        for (int64_t i = b; i < e; ++i) {
            const uint64_t f = (uint64_t)__ldg(feat_lo + i) | (feat_hi ? ((uint64_t)__ldg(feat_hi + i) << 32) : 0ull);
            const uint32_t v = hash64_dev(f + 1ull, (uint32_t)j + 1u);
            best = v < best ? v : best;
        }
        sig[row * (int64_t)H + j] = best;
    }
}

static int pick_R(int slabs) {
    // largest R in [4,8] wasting the fewest slabs in the last unit
    int best = 5, best_waste = 1 << 30;
    for (int R = 8; R >= 4; --R) {
        const int waste = ((slabs + R - 1) / R) * R - slabs;
        if (waste < best_waste) { best_waste = waste; best = R; }
    }
    return best;
}

#ifndef K1_VARIANT
#define K1_VARIANT 2
#endif

template <int V>
static void k1_dispatch(int R, unsigned grid, dim3 block, cudaStream_t stream, const int64_t* d_indptr,
                        const uint32_t* d_feat, int64_t n_rows, int64_t nnz, int64_t indptr0, int H, int units,
                        uint32_t* d_sig) {
#define K1_LAUNCH(RR) k1_signatures32_kernel<RR, V><<<grid, block, 0, stream>>>(d_indptr, d_feat, n_rows, nnz, indptr0, H, units, 1u, d_sig)
    switch (R) {
        case 4: K1_LAUNCH(4); break;
        case 5: K1_LAUNCH(5); break;
        case 6: K1_LAUNCH(6); break;
        case 7: K1_LAUNCH(7); break;
        default: K1_LAUNCH(8); break;
    }
#undef K1_LAUNCH
}

int launch_signatures32_variant(int variant, const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows,
                                int64_t nnz, int64_t indptr0, int H, uint32_t* d_sig, cudaStream_t stream) {
    if (n_rows <= 0) return 0;
    const int64_t total = n_rows * (int64_t)H;
    const int64_t want = (total + 1023) / 1024;
    const int fill_blocks = (int)(want < 148 * 8 ? want : 148 * 8);
    fill_u32_kernel<<<fill_blocks, 256, 0, stream>>>(d_sig, total, kModulus);
    if (nnz <= 0) return 1;
    const int slabs = (H + 31) / 32;
    const int R = pick_R(slabs);
    const int units = (slabs + R - 1) / R;
    int n_warps = units < K1_MAX_WARPS ? units : K1_MAX_WARPS;
    for (int w = K1_MAX_WARPS; w >= 4; --w)
        if (units % w == 0) { n_warps = w; break; }   // balanced split of units over warps
    const unsigned grid = (unsigned)((nnz + K1_CHUNK - 1) / K1_CHUNK);
    const dim3 block(32 * n_warps);
    if (variant < 0) {
        k1_dispatch<K1_VARIANT>(R, grid, block, stream, d_indptr, d_feat, n_rows, nnz, indptr0, H, units, d_sig);
        return 2;
    }
#ifdef K1_ALL_VARIANTS
    switch (variant) {
#define K1_V(VV) case VV: k1_dispatch<VV>(5, grid, block, stream, d_indptr, d_feat, n_rows, nnz, indptr0, H, (slabs + 4) / 5, d_sig); break;
        K1_V(0) K1_V(1) K1_V(2) K1_V(3) K1_V(4) K1_V(5) K1_V(6) K1_V(7)
        K1_V(8) K1_V(9) K1_V(10) K1_V(11) K1_V(12) K1_V(13) K1_V(14) K1_V(15)
        K1_V(16) K1_V(17) K1_V(18) K1_V(19)
#undef K1_V
        default: return -1;
    }
    return 2;
#else
    return -1;
#endif
}

int launch_signatures32(const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows, int64_t nnz,
                        int64_t indptr0, int H, uint32_t* d_sig, cudaStream_t stream) {
    return launch_signatures32_variant(-1, d_indptr, d_feat, n_rows, nnz, indptr0, H, d_sig, stream);
}

int launch_signatures64(const int64_t* d_indptr, const uint32_t* d_feat_lo, const uint32_t* d_feat_hi,
                        int64_t n_rows, int64_t nnz, int64_t indptr0, int H, uint32_t* d_sig,
                        cudaStream_t stream) {
    (void)nnz;
    if (n_rows <= 0) return 0;
    const int by = (H + 255) / 256;
    dim3 grid((unsigned)n_rows, (unsigned)by);
    k1_signatures64_kernel<<<grid, 256, 0, stream>>>(d_indptr, d_feat_lo, d_feat_hi, H, d_sig); (void)indptr0;
    return 1;
}

}  // namespace schic
