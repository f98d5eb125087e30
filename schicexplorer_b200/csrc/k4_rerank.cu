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

constexpr int K4_THREADS = 256;
constexpr int K4_SLOTS = 8192;            // hash table slots (32 KB keys + 32 KB values)
constexpr int K4_BUCKETS = K4_SLOTS / 4;  // a bucket = 4 consecutive keys = one LDS.128
constexpr int K4_BUCKET_BITS = 11;
constexpr int K4_PART = 2560;             // query non-zeros per table fill: load factor 0.31
constexpr uint32_t K4_EMPTY = 0xFFFFFFFFu;

__global__ void row_norms_kernel(const int64_t* __restrict__ indptr, const float* __restrict__ val, int64_t n_rows,
                                 double* __restrict__ norm2) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    double acc = 0.0;
    for (int64_t i = indptr[row] + lane; i < indptr[row + 1]; i += 32) {
        const double v = (double)__ldg(val + i);
        acc += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) norm2[row] = acc;
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
                 const int32_t* __restrict__ cand_nvalid, int kcand, int P, int k, int32_t* __restrict__ idx_out,
                 float* __restrict__ dist_out, int32_t* __restrict__ nvalid_out) {
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
    const int ncand = min(cand_nvalid[q], kcand);
    const int lane = threadIdx.x & 31;

    for (int c = threadIdx.x; c < ncand; c += blockDim.x) {
        const int32_t d = cand_idx[q * kcand + c];
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
    for (int c = threadIdx.x; c < P; c += blockDim.x) {
        unsigned long long key = ~0ull;
        if (c < ncand) {
            const int32_t d = cand[c];
            double d2 = nq2 + norm2[d] - 2.0 * dot[c];
            if (d2 < 0.0 || (int64_t)d == qrow) d2 = 0.0;
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
        idx_out[q * k + t] = oi;
        dist_out[q * k + t] = od;
    }
    if (threadIdx.x == 0) nvalid_out[q] = nv;
}

int launch_row_norms(const int64_t* d_indptr, const float* d_val, int64_t n_rows, int64_t indptr0,
                     double* d_norm2, cudaStream_t stream) {
    (void)indptr0;
    if (n_rows <= 0) return 0;
    const int64_t threads = n_rows * 32;
    row_norms_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(d_indptr, d_val, n_rows, d_norm2);
    return 1;
}

int launch_rerank(const int64_t* d_indptr, const uint32_t* d_feat, const float* d_val, int64_t indptr0,
                  const double* d_norm2, int64_t q_row0, int64_t nq, const int32_t* d_cand_idx,
                  const int32_t* d_cand_nvalid, int kcand, int k, int32_t* d_idx_out, float* d_dist_out,
                  int32_t* d_nvalid_out, cudaStream_t stream) {
    (void)indptr0;
    if (nq <= 0) return 0;
    const int P = next_pow2(kcand);
    const size_t smem = (size_t)K4_SLOTS * 8 + (size_t)P * 8 + (size_t)kcand * (8 + 8 + 8 + 4) + 16;
    if (smem > 220 * 1024) return -1;   // caller reports SCHIC_ERR_UNSUPPORTED
    cudaFuncSetAttribute(k4_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k4_rerank_kernel<<<(unsigned)nq, K4_THREADS, smem, stream>>>(d_indptr, d_feat, d_val, d_norm2, q_row0, d_cand_idx,
                                                                 d_cand_nvalid, kcand, P, k, d_idx_out, d_dist_out,
                                                                 d_nvalid_out);
    return 1;
}

}  // namespace schic
