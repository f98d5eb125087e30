// k5_graph.cu -- K5: kNN graph emit (CSR), sm_100a.
//
// Replaces the graph assembly of sparse_neighbors_search.MinHash.kneighbors_graph(mode='distance')
// (/root/reference/schicexplorer/scHicClusterMinHash.py:355; SURVEY 8(c) rule 10): CSR n x n_fitted,
// row c = its <= k (neighbour, distance) pairs sorted by column. indptr is an exclusive scan of the
// per-row valid counts (one CTA, decoupled tiles); each row is ordered by a shared-memory bitonic
// sort on packed (column, fp32 bits) keys and written with coalesced stores.
#include "common.cuh"
#include "sort.cuh"

namespace schic {

__global__ void __launch_bounds__(1024)
scan_nvalid_kernel(const int32_t* __restrict__ nvalid, int64_t n, int64_t* __restrict__ indptr) {
    __shared__ int scratch[33];
    int64_t running = 0;
    if (threadIdx.x == 0) indptr[0] = 0;
    for (int64_t base = 0; base < n; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const int v = i < n ? nvalid[i] : 0;
        int total;
        const int excl = block_exclusive_scan(v, scratch, &total);
        if (i < n) indptr[i + 1] = running + excl + v;
        running += total;
    }
}

__global__ void __launch_bounds__(128)
graph_emit_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dist, const int32_t* __restrict__ nvalid,
                  int k, int P, const int64_t* __restrict__ indptr, int32_t* __restrict__ indices,
                  float* __restrict__ data) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    const int64_t q = blockIdx.x;
    const int nv = nvalid[q];
    for (int t = threadIdx.x; t < P; t += blockDim.x) {
        unsigned long long key = ~0ull;
        if (t < nv) key = ((unsigned long long)(uint32_t)idx[q * k + t] << 32) | __float_as_uint(dist[q * k + t]);
        keys[t] = key;
    }
    bitonic_sort_shared(keys, P);
    const int64_t o = indptr[q];
    for (int t = threadIdx.x; t < nv; t += blockDim.x) {
        const unsigned long long key = keys[t];
        indices[o + t] = (int32_t)(uint32_t)(key >> 32);
        data[o + t] = __uint_as_float((uint32_t)(key & 0xFFFFFFFFull));
    }
}

int launch_scan_nvalid(const int32_t* d_nvalid, int64_t nq, int64_t* d_indptr, cudaStream_t stream) {
    scan_nvalid_kernel<<<1, 1024, 0, stream>>>(d_nvalid, nq, d_indptr);
    return 1;
}

int launch_graph_emit(const int32_t* d_idx, const float* d_dist, const int32_t* d_nvalid, int64_t nq, int k,
                      int64_t* d_indptr, int32_t* d_indices, float* d_data, cudaStream_t stream) {
    if (nq <= 0) return 0;
    const int P = next_pow2(k);
    const size_t smem = (size_t)P * 8;
    if (smem > 200 * 1024) return -1;
    scan_nvalid_kernel<<<1, 1024, 0, stream>>>(d_nvalid, nq, d_indptr);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(graph_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    graph_emit_kernel<<<(unsigned)nq, 128, smem, stream>>>(d_idx, d_dist, d_nvalid, k, P, d_indptr, d_indices, d_data);
    return 2;
}

}  // namespace schic
