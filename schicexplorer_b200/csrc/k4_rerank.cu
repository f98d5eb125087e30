// k4_rerank.cu -- K4: exact euclidean rerank of the MinHash candidates, sm_100a.
//
// Replaces the fast=False branch of sparse_neighbors_search.MinHash.kneighbors (selected by
// `-em`, /root/reference/schicexplorer/scHicClusterMinHash.py:66-68 -> fast=args.euclideanModeMinHash
// at :348/:403; SURVEY 8(c) rule 9): for every query cell and each of its <= k*excess_factor
// candidates, the euclidean distance between the two raw sparse rows, then (distance asc,
// index asc) order, keep k.
//
// The reference merge-joins two sorted index lists per pair. Here one CTA owns one query row:
//   * the query row is scattered once into an open-addressing hash table in shared memory
//     (feature id -> value), in feature-range parts when it does not fit, and is then reused for
//     all of the query's candidates;
//   * each warp streams one candidate row from HBM/L2 with coalesced loads and probes the table;
//     only x.y is accumulated (fp64), ||x-y||^2 = ||x||^2 + ||y||^2 - 2 x.y with fp64 row norms
//     computed once per fit. No tensor cores: nothing here is a dense contraction.
//   * the <= kcand distances are ordered with an in-shared-memory bitonic sort on packed
//     (fp32 distance bits, index) keys.
// Traffic per pair is 8 bytes per candidate non-zero (index + value), the HBM roofline of this
// stage; the query side costs nothing after the first touch.
#include "common.cuh"
#include "sort.cuh"

namespace schic {

// Which slice of every query's candidate list a launch works on and where its (sorted) results go: lists longer than
// the shared-memory limit are reranked in chunks and merged by k4_merge_rows_kernel.
struct K4Chunk {
    int cand_stride;   // row stride of cand_idx
    int cand_off;      // first candidate column of this launch
    int out_stride;    // row stride of idx_out / dist_out
    int out_off;       // first output column
    int write_nvalid;  // 1: this launch produces the final lists (also writes nvalid_out)
    int n_all;         // cand_idx == nullptr (brute force): candidate c of every query is database row cand_off + c, n_all rows in all
};

constexpr int K4_THREADS = 256;
constexpr int K4_SLOTS = 8192;            // hash table slots (32 KB keys + 32 KB values)
constexpr int K4_BUCKETS = K4_SLOTS / 4;  // a bucket = 4 consecutive keys = one LDS.128
constexpr int K4_BUCKET_BITS = 11;
constexpr int K4_PART = 2560;             // query non-zeros per table fill: load factor 0.31
constexpr uint32_t K4_EMPTY = 0xFFFFFFFFu;

// uint16 / uint8 counts (as uploaded) -> float32 values
template <typename T>
__global__ void widen_counts_kernel(const T* __restrict__ src, int64_t n, float* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (float)src[i];
}

int launch_widen_counts(const void* d_src, int bits, int64_t n, float* d_dst, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (bits == 16) widen_counts_kernel<uint16_t><<<148 * 8, 256, 0, stream>>>((const uint16_t*)d_src, n, d_dst);
    else widen_counts_kernel<uint8_t><<<148 * 8, 256, 0, stream>>>((const uint8_t*)d_src, n, d_dst);
    return 1;
}

// ||x - y||^2 of two sorted sparse rows by a plain two-pointer merge in fp64, term by term (the way the oracle and
// the reference's CPU path do it). The kernels below use ||x||^2 + ||y||^2 - 2 x.y, which is exact for integer
// counts but cancels for (near-)identical rows with non-integer values (fp32 products: absolute error ~1e-7 of the
// norms); a pair whose distance comes out below 10 % of the norm scale is redone with this function by the thread
// that owns it -- rare (duplicates, near-duplicates), and then identical rows give exactly 0.
__device__ __noinline__ double exact_sqdist_merge(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat,
                                                  const float* __restrict__ val, int64_t ra, int64_t rb) {
    int64_t i = indptr[ra], j = indptr[rb];
    const int64_t ie = indptr[ra + 1], je = indptr[rb + 1];
    double acc = 0.0;
    while (i < ie && j < je) {
        const uint32_t fa = __ldg(feat + i), fb = __ldg(feat + j);
        double t;
        if (fa == fb) { t = (double)__ldg(val + i) - (double)__ldg(val + j); ++i; ++j; }
        else if (fa < fb) { t = (double)__ldg(val + i); ++i; }
        else { t = (double)__ldg(val + j); ++j; }
        acc += t * t;
    }
    for (; i < ie; ++i) { const double t = (double)__ldg(val + i); acc += t * t; }
    for (; j < je; ++j) { const double t = (double)__ldg(val + j); acc += t * t; }
    return acc;
}

// d^2 from the norms and the dot product; pairs that cancel are redone exactly (see exact_sqdist_merge)
__device__ __forceinline__ double k4_sqdist(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat,
                                            const float* __restrict__ val, double nq2, double nd2, double xy,
                                            int64_t qrow, int64_t d) {
    if (d == qrow) return 0.0;
    double d2 = nq2 + nd2 - 2.0 * xy;
    if (d2 < 1e-2 * (nq2 + nd2)) d2 = exact_sqdist_merge(indptr, feat, val, qrow, d);
    return d2 < 0.0 ? 0.0 : d2;
}

// the counts kernel (k4c_rerank.cu), launched before these kernels, has done the work when the prep flags say the data
// is packable; same predicate as k4c_usable
__device__ __forceinline__ bool k4_skipped(const uint32_t* __restrict__ skip_flags, uint32_t esc_limit) {
    return skip_flags != nullptr && skip_flags[0] == 0u && skip_flags[1] <= esc_limit;
}

__device__ __forceinline__ uint32_t k4_bucket(uint32_t f) { return (f * 2654435761u) >> (32 - K4_BUCKET_BITS); }

// first position in [b, e) whose feature is >= key (plain binary search, one thread)
__device__ __forceinline__ int64_t lower_bound_feat(const uint32_t* __restrict__ feat, int64_t b, int64_t e, uint32_t key) {
    while (b < e) {
        const int64_t mid = (b + e) >> 1;
        if (__ldg(feat + mid) < key) b = mid + 1; else e = mid;
    }
    return b;
}

// One probe of one element: fetch the 4 keys of bucket b, report hit slot / miss / "bucket full, go on".
struct K4Probe { bool found; bool more; int slot; };
__device__ __forceinline__ K4Probe k4_probe(const uint32_t* __restrict__ tkey, uint32_t b, uint32_t f) {
    const uint4 k = *reinterpret_cast<const uint4*>(tkey + 4 * b);
    K4Probe r;
    const bool m0 = k.x == f, m1 = k.y == f, m2 = k.z == f, m3 = k.w == f;
    r.found = m0 || m1 || m2 || m3;
    r.slot = (int)(4 * b) + (m1 ? 1 : (m2 ? 2 : (m3 ? 3 : 0)));
    r.more = !r.found && k.w != K4_EMPTY;    // slots fill in order, so a free last slot ends the search
    return r;
}

__global__ void __launch_bounds__(K4_THREADS)
k4_rerank_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat, const float* __restrict__ val,
                 const double* __restrict__ norm2, int64_t q_row0, const int32_t* __restrict__ cand_idx,
                 const int32_t* __restrict__ cand_nvalid, int kcand, int P, int k, K4Chunk ck, int32_t* __restrict__ idx_out,
                 float* __restrict__ dist_out, int32_t* __restrict__ nvalid_out, const uint32_t* __restrict__ skip_flags,
                 uint32_t esc_limit) {
    if (k4_skipped(skip_flags, esc_limit)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* tkey = reinterpret_cast<uint32_t*>(smem_raw);              // [K4_SLOTS]
    float* tval = reinterpret_cast<float*>(tkey + K4_SLOTS);             // [K4_SLOTS]
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(tval + K4_SLOTS);   // [P]
    double* dot = reinterpret_cast<double*>(keys + P);                   // [kcand]
    int64_t* cursor = reinterpret_cast<int64_t*>(dot + kcand);           // [kcand] next unread y position
    int64_t* slice_end = cursor + kcand;                                 // [kcand] end of this part's slice
    int32_t* cand = reinterpret_cast<int32_t*>(slice_end + kcand);       // [kcand]
    __shared__ int s_next;

    const int64_t q = blockIdx.x;
    const int64_t qrow = q_row0 + q;
    const int ncand = max(0, min((cand_idx ? cand_nvalid[q] : ck.n_all) - ck.cand_off, kcand));
    const int lane = threadIdx.x & 31;

    for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
        const int32_t d = cand_idx ? cand_idx[q * ck.cand_stride + ck.cand_off + c] : ck.cand_off + c;
        cand[c] = d;
        dot[c] = 0.0;
        cursor[c] = indptr[d];
    }
    const int64_t xb = indptr[qrow], xe = indptr[qrow + 1];

    for (int64_t xs = xb; xs < xe; xs += K4_PART) {
        const int64_t xt = min(xs + (int64_t)K4_PART, xe);
        // exclusive upper feature bound of this part (features are strictly increasing in a row)
        const bool last = xt >= xe;
        const uint32_t hi = last ? 0u : __ldg(feat + xt);
        __syncthreads();   // previous part fully consumed (and cand/dot/cursor initialised)
        for (int i = threadIdx.x; i < K4_SLOTS / 4; i += blockDim.x)
            reinterpret_cast<uint4*>(tkey)[i] = make_uint4(K4_EMPTY, K4_EMPTY, K4_EMPTY, K4_EMPTY);
        if (threadIdx.x == 0) s_next = 0;
        __syncthreads();
        // scatter this part of the query row: first free slot of the home bucket, else the next bucket
        for (int64_t i = xs + threadIdx.x; i < xt; i += blockDim.x) {
            const uint32_t f = __ldg(feat + i);
            if (f == K4_EMPTY) continue;     // the one id that equals the empty marker: added out of band below
            uint32_t slot = 4 * k4_bucket(f);
            while (atomicCAS(&tkey[slot], K4_EMPTY, f) != K4_EMPTY) slot = (slot + 1) & (K4_SLOTS - 1);
            tval[slot] = __ldg(val + i);
        }
        // where each candidate's slice for this feature range ends: one binary search per candidate,
        // all candidates in parallel (their latency overlaps the scatter of the other warps)
        for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
            const int64_t ye = indptr[cand[c] + 1];
            slice_end[c] = last ? ye : lower_bound_feat(feat, cursor[c], ye, hi);
        }
        __syncthreads();
        // candidates are handed to warps dynamically (their slices differ in length)
        for (;;) {
            int c = 0;
            if (lane == 0) c = atomicAdd(&s_next, 1);
            c = __shfl_sync(0xffffffffu, c, 0);
            if (c >= ncand) break;
            const int64_t ys = cursor[c];
            const int len = (int)(slice_end[c] - ys);
            const uint32_t* __restrict__ yf = feat + ys;
            const float* __restrict__ yvp = val + ys;
            // Stream the slice, 4 independent elements per lane and iteration; the loads of the next
            // iteration are issued before the current one is probed (register double buffering).
            double acc = 0.0;
            uint32_t fn[4];
            float vn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = u * 32 + lane;
                fn[u] = i < len ? __ldg(yf + i) : K4_EMPTY;
                vn[u] = i < len ? __ldg(yvp + i) : 0.f;
            }
            for (int i0 = 0; i0 < len; i0 += 128) {
                uint32_t f[4];
                float yv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { f[u] = fn[u]; yv[u] = vn[u]; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 128 + u * 32 + lane;
                    fn[u] = i < len ? __ldg(yf + i) : K4_EMPTY;
                    vn[u] = i < len ? __ldg(yvp + i) : 0.f;
                }
                uint32_t b[4];
                K4Probe pr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {       // first probes: 4 independent LDS.128 in flight
                    b[u] = k4_bucket(f[u]);
                    pr[u] = k4_probe(tkey, b[u], f[u]);
                    if (f[u] == K4_EMPTY) { pr[u].found = false; pr[u].more = false; }
                }
                while (pr[0].more || pr[1].more || pr[2].more || pr[3].more) {   // rare: home bucket was full
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (pr[u].more) {
                            b[u] = (b[u] + 1) & (K4_BUCKETS - 1);
                            pr[u] = k4_probe(tkey, b[u], f[u]);
                        }
                    }
                }
                float prod[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) prod[u] = pr[u].found ? tval[pr[u].slot] * yv[u] : 0.f;
                // products of fp32 counts are exact in fp32 for integer data; the sum is kept in fp64
                acc += (double)prod[0] + (double)prod[1];
                acc += (double)prod[2] + (double)prod[3];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) dot[c] += acc;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < ncand; c += blockDim.x) cursor[c] = slice_end[c];
    }
    __syncthreads();

    // distances -> packed keys -> sort by (distance asc, index asc)
    const double nq2 = norm2[qrow];
    const bool x_has_max = xe > xb && __ldg(feat + xe - 1) == K4_EMPTY;
    for (int c = threadIdx.x; c < P; c += blockDim.x) {
        unsigned long long key = ~0ull;
        if (c < ncand) {
            const int32_t d = cand[c];
            double xy = dot[c];
            // id 2^32-1 doubles as the table's empty marker and is skipped by the table; rows are sorted, so
            // it can only be the last non-zero of a row: its product is added here
            if (x_has_max) {
                const int64_t yb = indptr[d], ye = indptr[d + 1];
                if (ye > yb && __ldg(feat + ye - 1) == K4_EMPTY)
                    xy += (double)__ldg(val + xe - 1) * (double)__ldg(val + ye - 1);
            }
            const double d2 = k4_sqdist(indptr, feat, val, nq2, norm2[d], xy, qrow, d);
            const float dist = (float)sqrt(d2);
            key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(uint32_t)d;
        }
        keys[c] = key;
    }
    bitonic_sort_shared(keys, P);
    const int nv = min(ncand, k);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        int32_t oi = -1;
        float od = __int_as_float(0x7f800000);
        if (t < nv) {
            const unsigned long long key = keys[t];
            oi = (int32_t)(uint32_t)(key & 0xFFFFFFFFull);
            od = __uint_as_float((uint32_t)(key >> 32));
        }
        idx_out[q * ck.out_stride + ck.out_off + t] = oi;
        dist_out[q * ck.out_stride + ck.out_off + t] = od;
    }
    if (threadIdx.x == 0 && ck.write_nvalid) nvalid_out[q] = nv;
}


// =================================================================================================
// Windowed variant (default when the feature space is small enough, i.e. every Hi-C resolution the
// CLI is used at): the hash probe above costs ~45 issued instructions per candidate non-zero (two
// LDS.128 probe rounds, 8 compares, selects) and the kernel sits on the ALU pipe, not on HBM. Here the
// feature axis is cut into fixed windows of 2^19 ids; for one window the query row becomes a
// direct-addressed table in shared memory -- one 32-bit entry per 16 feature ids holding a 16-bit
// presence mask and the rank of the entry's first key -- so a probe is one LDS, a bit test, a POPC and
// one LDS of the value: ~15 instructions, no probe sequence, no 128-bit bank conflicts.
// A per-row "window directory" (offset of each window's first non-zero, built once per fit by
// row_window_dir_kernel) replaces the per-candidate binary searches: a candidate's slice for window w
// is [dir[w], dir[w+1]). Non-zeros of a candidate that fall in windows where the query row is empty
// are never read.
// =================================================================================================
#ifndef K4W_THREADS_DEF
#define K4W_THREADS_DEF 1024
#endif
#ifndef K4W_UNROLL_DEF
#define K4W_UNROLL_DEF 8
#endif
constexpr int K4W_THREADS = K4W_THREADS_DEF;
#ifndef K4W_WIN_BITS_DEF
#define K4W_WIN_BITS_DEF 19
#endif
#ifndef K4W_CTAS_DEF
#define K4W_CTAS_DEF 1
#endif
constexpr int K4W_WIN_BITS = K4W_WIN_BITS_DEF;          // feature ids per window
constexpr int K4W_WORDS = 1 << (K4W_WIN_BITS - 4);      // table entries (128 KB: one CTA per SM)
constexpr int K4W_KCAP = 4096;                          // query non-zeros per table fill (values: 16 KB)
constexpr int K4W_UNROLL = K4W_UNROLL_DEF;              // candidate non-zeros per lane and iteration

__global__ void row_max_feature_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat,
                                       int64_t n_rows, uint32_t* __restrict__ out_max) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0;
    if (row < n_rows) {
        const int64_t b = indptr[row], e = indptr[row + 1];
        if (e > b) m = __ldg(feat + e - 1);    // rows are sorted: the last id is the largest
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out_max, m);
}

// Table entry layouts (E64 = false / true), both 4 feature ids per byte of shared memory:
//   32-bit entry per 16 ids: presence mask in the low half, rank of the entry's first key in the high half;
//   64-bit entry per 32 ids: {presence mask, BYTE offset of the first key's value} -- one LDS.64, and the
//   value address is a single LEA (rank * 4 + offset).
template <bool E64>
__device__ __forceinline__ float k4w_probe(const uint32_t* __restrict__ tab, const float* __restrict__ tval, uint32_t t,
                                           float yv, float acc) {
    if (E64) {
        const uint2 e = reinterpret_cast<const uint2*>(tab)[t >> 5];
        const uint32_t m = 1u << (t & 31u);
        const uint32_t off = e.y + 4u * (uint32_t)__popc(e.x & (m - 1u));
        const float tv = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(tval) + off);
        return (e.x & m) ? fmaf(tv, yv, acc) : acc;
    } else {
        const uint32_t e = tab[t >> 4];
        const uint32_t m = 1u << (t & 15u);
        const uint32_t idx = (e >> 16) + __popc(e & (m - 1u));
        const float tv = tval[idx];                     // always in bounds (tval is padded by 32)
        return (e & m) ? fmaf(tv, yv, acc) : acc;
    }
}

template <bool E64>
__global__ void __launch_bounds__(K4W_THREADS, K4W_CTAS_DEF)
k4_rerank_window_kernel(const int64_t* __restrict__ indptr, const uint32_t* __restrict__ feat,
                        const float* __restrict__ val, const double* __restrict__ norm2,
                        const uint32_t* __restrict__ dir, int n_windows, int dir_windows, int64_t q_row0,
                        const int32_t* __restrict__ cand_idx, const int32_t* __restrict__ cand_nvalid, int kcand, int P,
                        int k, K4Chunk ck, int32_t* __restrict__ idx_out, float* __restrict__ dist_out,
                        int32_t* __restrict__ nvalid_out, const uint32_t* __restrict__ skip_flags, uint32_t esc_limit) {
    if (k4_skipped(skip_flags, esc_limit)) return;
    // the directory is kept for the finer windows of the counts kernel: window w of this kernel starts at directory
    // entry min(w << DSH, dir_windows)
    constexpr int DSH = K4W_WIN_BITS - SCHIC_K4C_WIN_BITS;
    auto dent = [&](int w) { return min(w << DSH, dir_windows); };
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem_raw);                      // [K4W_WORDS]
    float* tval = reinterpret_cast<float*>(tab + K4W_WORDS);                    // [K4W_KCAP + 32]
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(tval + K4W_KCAP + 32);   // [P]
    double* dot = reinterpret_cast<double*>(keys + P);                          // [kcand]
    int64_t* sl_pos = reinterpret_cast<int64_t*>(dot + kcand);                  // [kcand] slice start (absolute)
    int64_t* cbeg = sl_pos + kcand;                                             // [kcand] row start of the candidate
    int32_t* sl_len = reinterpret_cast<int32_t*>(cbeg + kcand);                 // [kcand]
    int32_t* cand = sl_len + kcand;                                             // [kcand]
    uint32_t* qdir = reinterpret_cast<uint32_t*>(cand + kcand);                 // [n_windows + 1]
    __shared__ int s_next;

    const int64_t q = blockIdx.x;
    const int64_t qrow = q_row0 + q;
    const int ncand = max(0, min((cand_idx ? cand_nvalid[q] : ck.n_all) - ck.cand_off, kcand));
    const int lane = threadIdx.x & 31;
    const int64_t nw1 = (int64_t)dir_windows + 1;

    for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
        const int32_t d = cand_idx ? cand_idx[q * ck.cand_stride + ck.cand_off + c] : ck.cand_off + c;
        cand[c] = d;
        cbeg[c] = indptr[d];
        dot[c] = 0.0;
    }
    for (int w = threadIdx.x; w <= n_windows; w += blockDim.x) qdir[w] = dir[qrow * nw1 + dent(w)];
    for (int i = threadIdx.x; i < K4W_KCAP + 32; i += blockDim.x) tval[i] = 0.f;
    for (int i = threadIdx.x; i < K4W_WORDS / 4; i += blockDim.x)
        reinterpret_cast<uint4*>(tab)[i] = make_uint4(0u, 0u, 0u, 0u);
    const int64_t xb = indptr[qrow];
    __syncthreads();

    for (int w = 0; w < n_windows; ++w) {
        const uint32_t k0 = qdir[w], k1 = qdir[w + 1];
        if (k1 == k0) continue;                          // the query row is empty in this window
        const uint32_t lo = (uint32_t)w << K4W_WIN_BITS;
        for (uint32_t i0 = k0; i0 < k1; i0 += K4W_KCAP) {    // more than KCAP keys in one window: sub-parts
            const uint32_t i1 = min(i0 + (uint32_t)K4W_KCAP, k1);
            const bool first = i0 == k0, last = i1 == k1;
            // (the previous part was fully consumed before its un-build below, so the slice arrays are free)
            if (threadIdx.x == 0) s_next = 0;
            // candidate slices of this part: straight from the window directory; a binary search only
            // at the inner borders of sub-parts
            const uint32_t f_lo = first ? 0u : __ldg(feat + xb + i0);
            const uint32_t f_hi = last ? 0u : __ldg(feat + xb + i1);
            for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
                const int64_t d = cand[c];
                const int64_t cb = cbeg[c];
                const uint32_t* __restrict__ dc = dir + d * nw1;
                const int64_t wb = cb + __ldg(dc + dent(w)), we = cb + __ldg(dc + dent(w + 1));
                const int64_t ys = first ? wb : lower_bound_feat(feat, wb, we, f_lo);
                const int64_t ye = last ? we : lower_bound_feat(feat, wb, we, f_hi);
                sl_pos[c] = ys;
                sl_len[c] = (int)(ye - ys);
            }
            __syncthreads();                             // table un-built by everyone, slices visible
            for (uint32_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
                const uint32_t t = __ldg(feat + xb + i) - lo;
                const uint32_t tp = i == i0 ? 0xFFFFFFFFu : __ldg(feat + xb + i - 1) - lo;
                if (E64) {
                    const uint32_t word = t >> 5;
                    atomicOr(&tab[2 * word], 1u << (t & 31u));
                    if (i == i0 || (tp >> 5) != word) tab[2 * word + 1] = (i - i0) * 4u;   // sole writer of this half
                } else {
                    const uint32_t word = t >> 4;
                    uint32_t add = 1u << (t & 15u);
                    if (i == i0 || (tp >> 4) != word) add |= (i - i0) << 16;
                    atomicOr(&tab[word], add);
                }
                tval[i - i0] = __ldg(val + xb + i);
            }
            __syncthreads();
            {   // pull what the NEXT part's set-up will read (query keys/values, the candidates' directory
                // entries) into L2 now, so that the two set-up phases do not sit on DRAM latency
                uint32_t nk0 = 0, nk1 = 0;
                int nwin = w;
                if (!last) {
                    nk0 = i1; nk1 = min(i1 + (uint32_t)K4W_KCAP, k1);
                } else {
                    nwin = w + 1;
                    while (nwin < n_windows && qdir[nwin + 1] == qdir[nwin]) ++nwin;
                    if (nwin < n_windows) { nk0 = qdir[nwin]; nk1 = min(qdir[nwin + 1], nk0 + (uint32_t)K4W_KCAP); }
                }
                if (nk1 > nk0) {
                    const uint32_t pi = nk0 + 32u * threadIdx.x;
                    if (pi < nk1 + 32u) {
                        const int64_t at = xb + min(pi, nk1 - 1u);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(feat + at));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(val + at));
                    }
                    for (int c = threadIdx.x; c < ncand; c += blockDim.x)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(dir + (int64_t)cand[c] * nw1 + dent(nwin)));
                }
            }
            // Streaming: every warp walks a flattened sequence of (candidate, 256-non-zero block) pairs --
            // candidates are handed out dynamically -- with the loads of the next block (also across a
            // candidate boundary) in flight while the current one is probed.
            {
                constexpr int BLK = 32 * K4W_UNROLL;
                auto fetch = [&]() {                     // next candidate with a non-empty slice, or ncand
                    int c;
                    do {
                        c = 0;
                        if (lane == 0) c = atomicAdd(&s_next, 1);
                        c = __shfl_sync(0xffffffffu, c, 0);
                    } while (c < ncand && sl_len[c] <= 0);
                    return c;
                };
                uint32_t fn[K4W_UNROLL];
                float vn[K4W_UNROLL];
                auto load = [&](const uint32_t* __restrict__ yf, const float* __restrict__ yvp, int off, int len) {
                    if (off + BLK <= len) {              // warp-uniform: full block, no guards
#pragma unroll
                        for (int u = 0; u < K4W_UNROLL; ++u) { fn[u] = __ldg(yf + off + 32 * u); vn[u] = __ldg(yvp + off + 32 * u); }
                    } else {
#pragma unroll
                        for (int u = 0; u < K4W_UNROLL; ++u) {
                            const bool ok = off + 32 * u + lane < len;
                            fn[u] = ok ? __ldg(yf + off + 32 * u) : lo;      // t = 0 with value 0: contributes nothing
                            vn[u] = ok ? __ldg(yvp + off + 32 * u) : 0.f;
                        }
                    }
                };
                int c = fetch();
                int off = 0, len = 0;
                const uint32_t* __restrict__ yf = feat;
                const float* __restrict__ yvp = val;
                if (c < ncand) {
                    len = sl_len[c]; yf = feat + sl_pos[c] + lane; yvp = val + sl_pos[c] + lane;
                    load(yf, yvp, 0, len);
                }
                double acc = 0.0;
                while (c < ncand) {
                    uint32_t f[K4W_UNROLL];
                    float yv[K4W_UNROLL];
#pragma unroll
                    for (int u = 0; u < K4W_UNROLL; ++u) { f[u] = fn[u]; yv[u] = vn[u]; }
                    const int cur = c;
                    const int fill = len - off;          // non-zeros in the current block (may exceed BLK)
                    off += BLK;
                    const bool done = off >= len;        // last block of the current candidate
                    if (done) {
                        c = fetch();
                        off = 0;
                        if (c < ncand) { len = sl_len[c]; yf = feat + sl_pos[c] + lane; yvp = val + sl_pos[c] + lane; }
                    }
                    if (c < ncand) load(yf, yvp, off, len);
                    float a32 = 0.f;                     // exact for integer counts: 8 products < 2^24
                    if (fill >= BLK) {                   // full block: all probes in one basic block (8 independent chains)
#pragma unroll
                        for (int u = 0; u < K4W_UNROLL; ++u) a32 = k4w_probe<E64>(tab, tval, f[u] - lo, yv[u], a32);
                    } else {
                        // short last blocks (slices average ~1.7 blocks) skip the probe pairs that are all padding
#pragma unroll
                        for (int g = 0; g < K4W_UNROLL / 2; ++g) {
                            if (g == 0 || fill > 64 * g) {
                                a32 = k4w_probe<E64>(tab, tval, f[2 * g] - lo, yv[2 * g], a32);
                                a32 = k4w_probe<E64>(tab, tval, f[2 * g + 1] - lo, yv[2 * g + 1], a32);
                            }
                        }
                    }
                    acc += (double)a32;
                    if (done) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                        if (lane == 0) dot[cur] += acc;
                        acc = 0.0;
                    }
                }
            }
            // sparse clear: zero exactly the entries this part set (the table is all zero between parts)
            __syncthreads();
            for (uint32_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
                const uint32_t t = __ldg(feat + xb + i) - lo;
                if (E64) { tab[2 * (t >> 5)] = 0u; tab[2 * (t >> 5) + 1] = 0u; } else { tab[t >> 4] = 0u; }
            }
        }
    }
    __syncthreads();

    // distances -> packed keys -> sort by (distance asc, index asc)
    const double nq2 = norm2[qrow];
    for (int c = threadIdx.x; c < P; c += blockDim.x) {
        unsigned long long key = ~0ull;
        if (c < ncand) {
            const int32_t d = cand[c];
            const double d2 = k4_sqdist(indptr, feat, val, nq2, norm2[d], dot[c], qrow, d);
            const float dist = (float)sqrt(d2);
            key = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned long long)(uint32_t)d;
        }
        keys[c] = key;
    }
    bitonic_sort_shared(keys, P);
    const int nv = min(ncand, k);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        int32_t oi = -1;
        float od = __int_as_float(0x7f800000);
        if (t < nv) {
            const unsigned long long key = keys[t];
            oi = (int32_t)(uint32_t)(key & 0xFFFFFFFFull);
            od = __uint_as_float((uint32_t)(key >> 32));
        }
        idx_out[q * ck.out_stride + ck.out_off + t] = oi;
        dist_out[q * ck.out_stride + ck.out_off + t] = od;
    }
    if (threadIdx.x == 0 && ck.write_nvalid) nvalid_out[q] = nv;
}

int launch_row_max_feature(const int64_t* d_indptr, const uint32_t* d_feat, int64_t n_rows, uint32_t* d_out_max,
                           cudaStream_t stream) {
    if (cudaMemsetAsync(d_out_max, 0, sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (n_rows <= 0) return 0;
    row_max_feature_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, stream>>>(d_indptr, d_feat, n_rows, d_out_max);
    return 1;
}

// Merge of chunked rerank results: one CTA per query sorts the row's [part_stride] (distance, index) pairs (the best
// of every chunk, each chunk sorted and padded with (inf, -1)) by (distance, index) in shared memory and keeps the
// first k. cand_nvalid == nullptr: brute force over n_all rows. exclude_row0 >= 0: the query's own row
// (exclude_row0 + q) is dropped (sklearn's kneighbors_graph() convention).
__global__ void __launch_bounds__(1024)
k4_merge_rows_kernel(const int32_t* __restrict__ part_idx, const float* __restrict__ part_dist, int part_stride,
                     const int32_t* __restrict__ cand_nvalid, int kcand, int n_all, int64_t exclude_row0, int P, int k,
                     int32_t* __restrict__ idx_out, float* __restrict__ dist_out, int32_t* __restrict__ nvalid_out) {
    extern __shared__ unsigned long long mkeys[];
    const int64_t q = blockIdx.x;
    const int64_t self = exclude_row0 >= 0 ? exclude_row0 + q : -1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned long long key = ~0ull;
        if (i < part_stride) {
            const int32_t d = part_idx[q * part_stride + i];
            // distances are >= 0, so their float bit patterns order like unsigned integers
            if (d >= 0 && d != self)
                key = ((unsigned long long)__float_as_uint(part_dist[q * part_stride + i]) << 32) | (uint32_t)d;
        }
        mkeys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool up = (i & size) == 0;
                    const unsigned long long a = mkeys[i], b = mkeys[j];
                    if ((a > b) == up) { mkeys[i] = b; mkeys[j] = a; }
                }
            }
            __syncthreads();
        }
    }
    const int total = cand_nvalid ? min(cand_nvalid[q], kcand) : n_all - (self >= 0 && self < n_all ? 1 : 0);
    const int nv = min(total, k);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const unsigned long long key = mkeys[t < P ? t : P - 1];
        const bool ok = t < nv && t < P && key != ~0ull;
        idx_out[q * k + t] = ok ? (int32_t)(uint32_t)key : -1;
        dist_out[q * k + t] = ok ? __uint_as_float((uint32_t)(key >> 32)) : INFINITY;
    }
    if (threadIdx.x == 0) nvalid_out[q] = nv;
}

int launch_rerank_merge(const int32_t* d_part_idx, const float* d_part_dist, int part_stride,
                        const int32_t* d_cand_nvalid, int kcand, int n_all, int64_t exclude_row0, int64_t nq, int k,
                        int32_t* d_idx_out, float* d_dist_out, int32_t* d_nvalid_out, cudaStream_t stream) {
    if (nq <= 0) return 0;
    const int P = next_pow2(part_stride);
    const size_t smem = (size_t)P * 8;
    if (smem > 220 * 1024) return -1;
    cudaFuncSetAttribute(k4_merge_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k4_merge_rows_kernel<<<(unsigned)nq, 1024, smem, stream>>>(d_part_idx, d_part_dist, part_stride, d_cand_nvalid, kcand,
                                                              n_all, exclude_row0, P, k, d_idx_out, d_dist_out,
                                                              d_nvalid_out);
    return 1;
}

// One launch over candidate columns [cand_off, cand_off + kcand) of lists with row stride cand_stride; results to
// output columns [out_off, out_off + k) of arrays with row stride out_stride. final != 0: nvalid_out is written.
// d_cand_idx == nullptr: brute force, the candidates of every query are the database rows [0, n_all).
// allow_hash == 0: only the windowed kernel may be used (returns -1 if it does not fit or there is no directory).
int launch_rerank(const int64_t* d_indptr, const uint32_t* d_feat, const float* d_val, int64_t indptr0,
                  const double* d_norm2, const uint32_t* d_dir, int n_windows, int window_mode, int64_t q_row0,
                  int64_t nq, const int32_t* d_cand_idx, const int32_t* d_cand_nvalid, int cand_stride, int cand_off,
                  int kcand, int k, int32_t* d_idx_out, float* d_dist_out, int out_stride, int out_off, int final,
                  int32_t* d_nvalid_out, int n_all, int allow_hash, const uint32_t* d_skip_flags, uint32_t esc_limit,
                  cudaStream_t stream) {
    (void)indptr0;
    if (nq <= 0) return 0;
    const K4Chunk ck{cand_stride, cand_off, out_stride, out_off, final, n_all};
    const int P = next_pow2(kcand);
    static_assert(K4W_WIN_BITS >= SCHIC_K4C_WIN_BITS, "the directory's windows must not be coarser than this kernel's");
    const int dir_windows = n_windows;                                            // windows of the directory (2^16 ids)
    constexpr int DSH = K4W_WIN_BITS - SCHIC_K4C_WIN_BITS;
    n_windows = (dir_windows + (1 << DSH) - 1) >> DSH;                            // windows of this kernel (2^19 ids)
    if (d_dir && n_windows > 0) {
        const size_t smem = (size_t)K4W_WORDS * 4 + (size_t)(K4W_KCAP + 32) * 4 + (size_t)P * 8 +
                            (size_t)kcand * (8 + 8 + 8 + 4 + 4) + (size_t)(n_windows + 1) * 4 + 16;
        if (smem <= 220 * 1024) {
            auto kern = window_mode == 64 ? k4_rerank_window_kernel<true> : k4_rerank_window_kernel<false>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<(unsigned)nq, K4W_THREADS, smem, stream>>>(d_indptr, d_feat, d_val, d_norm2, d_dir, n_windows, dir_windows,
                                                              q_row0, d_cand_idx, d_cand_nvalid, kcand, P, k, ck, d_idx_out,
                                                              d_dist_out, d_nvalid_out, d_skip_flags, esc_limit);
            return 1;
        }
    }
    if (!allow_hash) return -1;         // the caller prefers the windowed kernel on chunks over this one on the whole list
    const size_t smem = (size_t)K4_SLOTS * 8 + (size_t)P * 8 + (size_t)kcand * (8 + 8 + 8 + 4) + 16;
    if (smem > 220 * 1024) return -1;   // caller chunks the candidate lists
    cudaFuncSetAttribute(k4_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k4_rerank_kernel<<<(unsigned)nq, K4_THREADS, smem, stream>>>(d_indptr, d_feat, d_val, d_norm2, q_row0, d_cand_idx,
                                                                 d_cand_nvalid, kcand, P, k, ck, d_idx_out, d_dist_out,
                                                                 d_nvalid_out, d_skip_flags, esc_limit);
    return 1;
}

}  // namespace schic
