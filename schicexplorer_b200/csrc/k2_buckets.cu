// k2_buckets.cu -- K2: max_bin_size pruning of oversized hash buckets, sm_100a.
//
// Upstream's inverse index ignores a bucket (all cells sharing one signature value in one hash table)
// once it holds >= max_bin_size cells (kwarg at /root/reference/schicexplorer/scHicClusterMinHash.py:403;
// SURVEY 8(c) rule 5). The reference learns bucket sizes from its hash maps; here the size of the
// bucket of (cell c, table j) is the column-wise twin of K3's count,
//     size[c][j] = #{ d : sig[d][j] == sig[c][j] },
// computed all-pairs with the same compare (ALU) + predicated FADD (FMA pipe) pairing, and the kernel
// writes a masked copy of the signature matrix in which members of oversized buckets become 0 -- a
// value K3 already treats as inadmissible -- so K3 itself needs no change. Runs only when
// n_fitted >= max_bin_size (never for the CLI default 100000 below 100 k cells).
#include "common.cuh"

namespace schic {

constexpr int K2_CELLS = 64;     // cells per CTA (8 per thread)
constexpr int K2_STAGE = 64;     // database rows per shared-memory stage

__global__ void __launch_bounds__(256)
k2_bucket_mask_kernel(const uint32_t* __restrict__ sig, int64_t n, int H, float max_bin, uint32_t* __restrict__ out) {
    __shared__ uint32_t stage[K2_STAGE][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    const int64_t c0 = (int64_t)blockIdx.y * K2_CELLS + warp * 8;
    const bool jok = j < H;

    uint32_t v[8];
    float cnt[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t c = c0 + a;
        v[a] = (jok && c < n) ? __ldg(sig + c * (int64_t)H + j) : 0xFFFFFFFFu;
        if (v[a] == 0u || v[a] == kModulus) v[a] = 0xFFFFFFFFu;   // inadmissible anyway: never counted
        cnt[a] = 0.f;
    }
    for (int64_t d0 = 0; d0 < n; d0 += K2_STAGE) {
        __syncthreads();
        for (int r = warp; r < K2_STAGE; r += 8) {
            const int64_t d = d0 + r;
            stage[r][lane] = (jok && d < n) ? __ldg(sig + d * (int64_t)H + j) : 0xFFFFFFFEu;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = 0; r < K2_STAGE; ++r) {
            const uint32_t w = stage[r][lane];
#pragma unroll
            for (int a = 0; a < 8; ++a)
                asm("{ .reg .pred p; setp.eq.u32 p, %1, %2; @p add.f32 %0, %0, 0f3F800000; }"
                    : "+f"(cnt[a]) : "r"(v[a]), "r"(w));
        }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t c = c0 + a;
        if (jok && c < n) {
            const uint32_t orig = __ldg(sig + c * (int64_t)H + j);
            out[c * (int64_t)H + j] = (cnt[a] >= max_bin) ? 0u : orig;
        }
    }
}

int launch_bucket_mask(const uint32_t* d_sig, int64_t n, int H, int64_t max_bin_size, uint32_t* d_out,
                       cudaStream_t stream) {
    if (n <= 0) return 0;
    dim3 grid((unsigned)((H + 31) / 32), (unsigned)((n + K2_CELLS - 1) / K2_CELLS));
    k2_bucket_mask_kernel<<<grid, 256, 0, stream>>>(d_sig, n, H, (float)max_bin_size, d_out);
    return 1;
}

}  // namespace schic
