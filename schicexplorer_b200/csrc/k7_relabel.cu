// k7_relabel.cu -- feature ids beyond 32 bits for the exact rerank, sm_100a.
//
// At 10 kb resolution a (cells x bins^2) matrix has 6.8e10 columns (/root/reference/docs/content/Example_analysis.rst:350-351),
// so feature ids need more than 32 bits. K1 hashes the full ids (width-64 HashSpec) or their low words (width 32, exact by
// definition of the hash), K3 only sees signatures -- but the rerank kernels compare ids, index tables by them and cut
// the id axis into windows. Distances depend only on which ids are EQUAL and on their ORDER inside a row (rows are sorted),
// so the rerank runs on the RANK of every id among the distinct ids of the fitted data: a monotone map into [0, #distinct),
// which fits 32 bits whenever the non-zeros do. One radix sort of (id, position) pairs per fit (CUB -- library code, off
// the hot path: only data sets with ids >= 2^32 ever come here), a flag + scan for the ranks, a scatter back.
#include <cub/cub.cuh>

#include "common.cuh"

namespace schic {

__global__ void relabel_keys_kernel(const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi, int64_t n,
                                    unsigned long long* __restrict__ keys, uint32_t* __restrict__ pos) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        keys[i] = ((unsigned long long)hi[i] << 32) | lo[i];
        pos[i] = (uint32_t)i;
    }
}

__global__ void relabel_flags_kernel(const unsigned long long* __restrict__ keys, int64_t n, uint32_t* __restrict__ flags) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        flags[i] = (i > 0 && keys[i] != keys[i - 1]) ? 1u : 0u;
}

__global__ void relabel_scatter_kernel(const uint32_t* __restrict__ rank, const uint32_t* __restrict__ pos, int64_t n,
                                       uint32_t* __restrict__ labels, uint32_t* __restrict__ max_label) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) labels[pos[i]] = rank[i];
    if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) *max_label = rank[n - 1];
}

// bytes of scratch for n ids
size_t relabel_scratch_bytes(int64_t n) {
    size_t sort_tmp = 0, scan_tmp = 0;
    cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, k, v, (int)std::min<int64_t>(n, INT32_MAX));
    cub::DeviceScan::InclusiveSum(nullptr, scan_tmp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)std::min<int64_t>(n, INT32_MAX));
    const size_t a = 256;
    auto up = [&](size_t x) { return (x + a - 1) / a * a; };
    return up((size_t)n * 8) * 2 + up((size_t)n * 4) * 3 + up(std::max(sort_tmp, scan_tmp)) + a;
}

// labels[i] = rank of id i = (hi[i] << 32 | lo[i]) among the distinct ids; *d_max_label = largest rank. n < 2^31.
int launch_relabel_ids(const uint32_t* d_lo, const uint32_t* d_hi, int64_t n, uint32_t* d_labels, uint32_t* d_max_label,
                       void* d_scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (n > INT32_MAX) return -1;
    const size_t a = 256;
    auto up = [&](size_t x) { return (x + a - 1) / a * a; };
    unsigned char* p = (unsigned char*)d_scratch;
    unsigned long long* k0 = (unsigned long long*)p; p += up((size_t)n * 8);
    unsigned long long* k1 = (unsigned long long*)p; p += up((size_t)n * 8);
    uint32_t* v0 = (uint32_t*)p; p += up((size_t)n * 4);
    uint32_t* v1 = (uint32_t*)p; p += up((size_t)n * 4);
    uint32_t* fl = (uint32_t*)p; p += up((size_t)n * 4);
    size_t tmp_bytes = scratch_bytes - (size_t)(p - (unsigned char*)d_scratch);
    relabel_keys_kernel<<<148 * 8, 256, 0, stream>>>(d_lo, d_hi, n, k0, v0);
    cub::DoubleBuffer<unsigned long long> kb(k0, k1);
    cub::DoubleBuffer<uint32_t> vb(v0, v1);
    if (cub::DeviceRadixSort::SortPairs(p, tmp_bytes, kb, vb, (int)n, 0, 64, stream) != cudaSuccess) return -1;
    relabel_flags_kernel<<<148 * 8, 256, 0, stream>>>(kb.Current(), n, fl);
    uint32_t* rank = vb.Alternate();                      // the other value buffer is free now
    if (cub::DeviceScan::InclusiveSum(p, tmp_bytes, fl, rank, (int)n, stream) != cudaSuccess) return -1;
    relabel_scatter_kernel<<<148 * 8, 256, 0, stream>>>(rank, vb.Current(), n, d_labels, d_max_label);
    return 6;
}

}  // namespace schic
