// sort.cuh -- block-wide helpers shared by the select / rerank / graph kernels.
#pragma once
#include <stdint.h>

namespace schic {

// In-place ascending bitonic sort of P (power of two) keys in shared memory by the whole CTA.
// Callers pad with the maximum key. Ends with a __syncthreads().
template <typename T>
__device__ __forceinline__ void bitonic_sort_shared(T* keys, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));   // index with bit `stride` cleared
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const T a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncthreads();
}

__host__ __device__ inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Exclusive prefix sum over the CTA of one int per thread (blockDim.x <= 1024, multiple of 32).
// `warp_sums` is shared scratch of >= 33 ints. Returns the exclusive prefix; *total gets the sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();   // protect warp_sums from the previous use
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < n_warps ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_sums[lane] = wi - w;          // exclusive warp offsets
        if (lane == 31) warp_sums[32] = wi; // grand total
    }
    __syncthreads();
    *total = warp_sums[32];
    return warp_sums[warp] + incl - v;
}

}  // namespace schic
