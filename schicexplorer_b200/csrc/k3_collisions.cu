// k3_collisions.cu -- K3: hash-bucket collision counts, sm_100a.
//
// Replaces the inverse-index lookup + per-query collision counting of
// sparse_neighbors_search.MinHash.kneighbors (reached through kneighbors_graph at
// /root/reference/schicexplorer/scHicClusterMinHash.py:355 and inside MinHashClustering.fit :406):
// two cells collide in table j when their signature values are equal (SURVEY 8(c) rules 5-6),
//     cnt[q][d] = #{ j : sigQ[q][j] == sigD[d][j],  value not in {0, M} }.
//
// The reference walks hash-map buckets per query; on the GPU the same counts are an all-pairs
// "equality GEMM" over the [n, H] signature matrix: CUDA cores only (compare + predicated add),
// 128x128 output tile per CTA, 8x8 register tile per thread, signature tiles staged transposed
// in shared memory so every operand fetch is a conflict-free LDS.128. Non-admissible values
// (0, the sentinel M, padding) are rewritten while staging to two patterns that can never be
// equal, so the inner loop has no validity test.
#include "common.cuh"

namespace schic {

constexpr int K3_TILE = 128;      // rows of Q and of D per CTA
constexpr int K3_JB = 32;         // hash functions per staging step
constexpr int K3_LD = K3_TILE + 4;  // padded leading dimension (keeps 16-byte alignment)
constexpr int K3_THREADS = 256;

// SYM: queries == database (same matrix, all rows in one launch). Counts are symmetric, so only tiles on
// or above the diagonal are computed and each off-diagonal tile is also stored transposed: half the compares.
template <bool SYM>
__global__ void __launch_bounds__(K3_THREADS, 2)
k3_collision_kernel(const uint32_t* __restrict__ sig_q, int64_t nq, const uint32_t* __restrict__ sig_d,
                    int64_t nd, int H, uint16_t* __restrict__ cnt, int64_t ld_out) {
    __shared__ __align__(16) uint32_t Qs[K3_JB][K3_LD];
    __shared__ __align__(16) uint32_t Ds[K3_JB][K3_LD];

    if (SYM && blockIdx.x < blockIdx.y) return;
    const int64_t q0 = (int64_t)blockIdx.y * K3_TILE;
    const int64_t d0 = (int64_t)blockIdx.x * K3_TILE;
    const int tx = threadIdx.x & 15;   // 8 database columns tx*8..
    const int ty = threadIdx.x >> 4;   // 8 query rows ty*8..
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // Counts are kept as fp32 (exact up to 2^24 >> H): the compare issues on the ALU pipe (ISETP)
    // and the predicated increment on the FMA pipe (FADD), so the two halves of every
    // compare-and-count pair dual-issue instead of queueing on one pipe.
    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

    for (int j0 = 0; j0 < H; j0 += K3_JB) {
        // stage: lane <-> hash (coalesced 128-byte row reads), 8 warps x 16 rows each, transposed store
        const int j = j0 + lane;
#pragma unroll 4
        for (int r = warp; r < K3_TILE; r += K3_THREADS / 32) {
            uint32_t vq = 0xFFFFFFFFu, vd = 0xFFFFFFFEu;
            if (j < H) {
                if (q0 + r < nq) {
                    const uint32_t v = __ldg(sig_q + (q0 + r) * (int64_t)H + j);
                    if (v != 0u && v != kModulus) vq = v;
                }
                if (d0 + r < nd) {
                    const uint32_t v = __ldg(sig_d + (d0 + r) * (int64_t)H + j);
                    if (v != 0u && v != kModulus) vd = v;
                }
            }
            Qs[lane][r] = vq;
            Ds[lane][r] = vd;
        }
        __syncthreads();
#pragma unroll 2
        for (int jj = 0; jj < K3_JB; ++jj) {
            const uint4 qa = *reinterpret_cast<const uint4*>(&Qs[jj][ty * 8]);
            const uint4 qb = *reinterpret_cast<const uint4*>(&Qs[jj][ty * 8 + 4]);
            const uint4 da = *reinterpret_cast<const uint4*>(&Ds[jj][tx * 8]);
            const uint4 db = *reinterpret_cast<const uint4*>(&Ds[jj][tx * 8 + 4]);
            const uint32_t q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const uint32_t d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    asm("{ .reg .pred p; setp.eq.u32 p, %1, %2; @p add.f32 %0, %0, 0f3F800000; }"
                        : "+f"(acc[a][b]) : "r"(q[a]), "r"(d[b]));
        }
        __syncthreads();
    }

    // store: 8 uint16 = 16 bytes per row when aligned
    const bool vec_ok = ((ld_out & 7) == 0) && ((((uintptr_t)cnt) & 15) == 0) && (d0 + tx * 8 + 8 <= nd);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t q = q0 + ty * 8 + a;
        if (q >= nq) continue;
        uint16_t* dst = cnt + q * ld_out + d0 + tx * 8;
        if (vec_ok) {
            uint4 o;
            o.x = (uint32_t)acc[a][0] | ((uint32_t)acc[a][1] << 16);
            o.y = (uint32_t)acc[a][2] | ((uint32_t)acc[a][3] << 16);
            o.z = (uint32_t)acc[a][4] | ((uint32_t)acc[a][5] << 16);
            o.w = (uint32_t)acc[a][6] | ((uint32_t)acc[a][7] << 16);
            *reinterpret_cast<uint4*>(dst) = o;
        } else {
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if (d0 + tx * 8 + b < nd) dst[b] = (uint16_t)acc[a][b];
        }
    }
    if (SYM && blockIdx.x > blockIdx.y) {
        // mirrored tile: row d0 + tx*8 + b, columns q0 + ty*8 .. +7 (16 bytes; 16 threads of equal tx cover 256 B)
        const bool tvec_ok = ((ld_out & 7) == 0) && ((((uintptr_t)cnt) & 15) == 0) && (q0 + ty * 8 + 8 <= nq);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int64_t d = d0 + tx * 8 + b;
            if (d >= nd) continue;
            uint16_t* dst = cnt + d * ld_out + q0 + ty * 8;
            if (tvec_ok) {
                uint4 o;
                o.x = (uint32_t)acc[0][b] | ((uint32_t)acc[1][b] << 16);
                o.y = (uint32_t)acc[2][b] | ((uint32_t)acc[3][b] << 16);
                o.z = (uint32_t)acc[4][b] | ((uint32_t)acc[5][b] << 16);
                o.w = (uint32_t)acc[6][b] | ((uint32_t)acc[7][b] << 16);
                *reinterpret_cast<uint4*>(dst) = o;
            } else {
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (q0 + ty * 8 + a < nq) dst[a] = (uint16_t)acc[a][b];
            }
        }
    }
}

int launch_collision_counts(const uint32_t* d_sig_q, int64_t nq, const uint32_t* d_sig_db, int64_t ndb,
                            int H, uint16_t* d_cnt, int64_t ld_out, cudaStream_t stream) {
    if (nq <= 0 || ndb <= 0) return 0;
    dim3 grid((unsigned)((ndb + K3_TILE - 1) / K3_TILE), (unsigned)((nq + K3_TILE - 1) / K3_TILE));
    if (d_sig_q == d_sig_db && nq == ndb)
        k3_collision_kernel<true><<<grid, K3_THREADS, 0, stream>>>(d_sig_q, nq, d_sig_db, ndb, H, d_cnt, ld_out);
    else
        k3_collision_kernel<false><<<grid, K3_THREADS, 0, stream>>>(d_sig_q, nq, d_sig_db, ndb, H, d_cnt, ld_out);
    return 1;
}

}  // namespace schic
