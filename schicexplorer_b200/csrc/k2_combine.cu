// k2_combine.cu -- K2: optional block combine ("shingle") of the MinHash minima, sm_100a.
//
// Upstream can fold the minima of `shingle_size` consecutive hash functions into one key, which leaves H / s hash
// tables (kwarg shingle_size at /root/reference/schicexplorer/scHicClusterMinHash.py:348, :403; whether it is
// active depends on a separate upstream switch the reference does not pass -- SURVEY 8(c) rule 4, decision point D2,
// default OFF). With schic_mh_params.shingle != 0 this stage runs behind K1:
//     key = sig[b*s];  key = mix((sig[b*s+t] + 1) * (key + 1)) mod (2^31 - 1)  for t = 1 .. s-1
// where mix is the hash family's own mix (Thomas Wang's, evaluated in uint32 or uint64 per HashSpec width). Rows
// without non-zeros keep the sentinel M in every table. One thread per (row, table); the raw matrix is read once.
#include "common.cuh"

namespace schic {

__device__ __forceinline__ uint32_t k2_wang32(uint32_t k) {
    k = ~k + (k << 15);
    k ^= k >> 12;
    k += k << 2;
    k ^= k >> 4;
    k *= 2057u;
    k ^= k >> 16;
    return k;
}
__device__ __forceinline__ uint32_t k2_wang64_mod(uint64_t k) {
    k = ~k + (k << 15);
    k ^= k >> 12;
    k += k << 2;
    k ^= k >> 4;
    k *= 2057ull;
    k ^= k >> 16;
    return (uint32_t)(k % (uint64_t)kModulus);
}

template <bool W64>
__global__ void __launch_bounds__(256)
k2_block_combine_kernel(const uint32_t* __restrict__ raw, int64_t n, int H_raw, int block, int T, uint32_t* __restrict__ keys) {
    const int64_t x = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (x >= n * (int64_t)T) return;
    const int64_t r = x / T;
    const int b = (int)(x - r * T);
    const uint32_t* __restrict__ s = raw + r * (int64_t)H_raw + (int64_t)b * block;
    uint32_t key = __ldg(s);
    if (__ldg(raw + r * (int64_t)H_raw) == kModulus) {        // empty row: every minimum is the sentinel
        keys[x] = kModulus;
        return;
    }
    for (int t = 1; t < block; ++t) {
        const uint32_t v = __ldg(s + t);
        key = W64 ? k2_wang64_mod(((uint64_t)v + 1ull) * ((uint64_t)key + 1ull)) : k2_wang32((v + 1u) * (key + 1u)) % kModulus;
    }
    keys[x] = key;
}

int launch_block_combine(int width, const uint32_t* d_raw, int64_t n, int H_raw, int block, uint32_t* d_keys, cudaStream_t stream) {
    const int T = H_raw / block;
    if (n <= 0 || T <= 0) return 0;
    const int64_t total = n * (int64_t)T;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (width == 64) k2_block_combine_kernel<true><<<grid, 256, 0, stream>>>(d_raw, n, H_raw, block, T, d_keys);
    else k2_block_combine_kernel<false><<<grid, 256, 0, stream>>>(d_raw, n, H_raw, block, T, d_keys);
    return 1;
}

}  // namespace schic
