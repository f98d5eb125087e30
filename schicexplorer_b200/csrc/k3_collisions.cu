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
#include <cuda.h>
#include <cuda_fp16.h>

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



// =================================================================================================
// Packed variant: two hash functions per compare instruction.
//
// The kernel above spends one ISETP (ALU pipe) + one FADD (FMA pipe) per signature compare; the ALU pipe
// is its roof (64 compares / clk / SM). Equality does not need the 31-bit values, only "same value in
// the same column": per hash function the values of the fitted cells are renamed to 16-bit ids through
// an open-addressing table in global memory (id = slot index, so equal values <-> equal ids, exactly),
// and the ids are written as *fp16 bit patterns of normal numbers* (never zero, never NaN), two hash
// functions per 32-bit word. Then one HSET2.BF.EQ compares two hash functions at once (1.0h / 0.0h per
// half) and one HADD2 accumulates both counts in a half2 register (exact: counts <= H/2 <= 2047).
// Inadmissible values (0, M), padding rows and the padding column of an odd H are NaN: never equal.
// Limits: H <= 4094 (fp16 count range), n <= K3H_MAX_ROWS (ids fit 61 440 normal fp16 patterns).
// =================================================================================================
constexpr int K3H_MAX_IDS = 61440;          // finite, non-zero, normal fp16 bit patterns: 2 * 30 * 1024
constexpr int K3H_MAX_ROWS = 57344;         // load factor of the renaming table stays <= 0.94 in the worst case
constexpr uint16_t K3H_NAN = 0x7E00u;

__host__ __device__ inline int k3h_table_capacity(int64_t n) {
    int64_t c = (n * 5) / 2 + 1024;
    return (int)(c > K3H_MAX_IDS ? K3H_MAX_IDS : c);
}

// table: [H][cap] uint32, zero-initialised (0 is inadmissible, so it can mean "empty")
__global__ void __launch_bounds__(256)
k3h_rename_kernel(const uint32_t* __restrict__ sig, int64_t n, int H, int Hp, int cap, uint32_t* __restrict__ table,
                  uint32_t* __restrict__ sizes /* nullable: [H][cap] bucket sizes (zeroed) */, uint16_t* __restrict__ sig16) {
    const int64_t total = n * (int64_t)Hp;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int64_t row = e / Hp;
        const int j = (int)(e - row * Hp);
        uint16_t out = K3H_NAN;
        if (j < H) {
            const uint32_t v = __ldg(sig + row * (int64_t)H + j);
            if (v != 0u && v != kModulus) {
                uint32_t* __restrict__ t = table + (int64_t)j * cap;
                uint32_t slot = (uint32_t)(((uint64_t)(v * 2654435761u) * (uint64_t)cap) >> 32);
                for (;;) {
                    const uint32_t old = atomicCAS(t + slot, 0u, v);
                    if (old == 0u || old == v) break;
                    slot = slot + 1u == (uint32_t)cap ? 0u : slot + 1u;
                }
                if (sizes) atomicAdd(sizes + (int64_t)j * cap + slot, 1u);      // the slot IS the bucket of this value
                out = slot < 30720u ? (uint16_t)(0x0400u + slot) : (uint16_t)(0x8400u + (slot - 30720u));
            }
        }
        sig16[e] = out;
    }
}

// max_bin_size (SURVEY 8(c) rule 5): members of buckets with >= max_bin cells never collide -> NaN. The bucket sizes
// come for free from the renaming (one counter per table slot): O(n H) instead of K2's all-pairs O(n^2 H).
__global__ void __launch_bounds__(256)
k3h_prune_kernel(uint16_t* __restrict__ sig16, int64_t n, int H, int Hp, int cap, const uint32_t* __restrict__ sizes,
                 uint32_t max_bin) {
    const int64_t total = n * (int64_t)Hp;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int j = (int)(e % Hp);
        const uint32_t p = sig16[e];
        if (j >= H || p == K3H_NAN) continue;
        const uint32_t slot = p < 0x8000u ? p - 0x0400u : (p - 0x8400u) + 30720u;
        if (__ldg(sizes + (int64_t)j * cap + slot) >= max_bin) sig16[e] = K3H_NAN;
    }
}

// MODE 0: plain block of counts. MODE 1 (SYM): queries == database, only tiles on or above the diagonal are computed and
// mirrored into the same matrix. MODE 2: plain block, every tile ALSO stored transposed into cnt_t [nd, ld_t] -- the
// block a peer rank needs when the (rank x rank) blocks of a sharded symmetric count matrix are computed only once.
template <int MODE>
__global__ void __launch_bounds__(K3_THREADS, 2)
k3h_collision_kernel(const uint32_t* __restrict__ q16, int64_t nq, const uint32_t* __restrict__ d16, int64_t nd, int Hw,
                     uint16_t* __restrict__ cnt, int64_t ld_out, uint16_t* __restrict__ cnt_t, int64_t ld_t) {
    constexpr bool SYM = MODE == 1;
    __shared__ __align__(16) uint32_t Qs[K3_JB][K3_LD];
    __shared__ __align__(16) uint32_t Ds[K3_JB][K3_LD];

    if (SYM && blockIdx.x < blockIdx.y) return;
    const int64_t q0 = (int64_t)blockIdx.y * K3_TILE;
    const int64_t d0 = (int64_t)blockIdx.x * K3_TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t NAN2 = 0x7E007E00u;

    uint32_t acc[8][8];                         // half2 counters: even / odd hash functions
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0u;

    for (int j0 = 0; j0 < Hw; j0 += K3_JB) {
        const int j = j0 + lane;
#pragma unroll 4
        for (int r = warp; r < K3_TILE; r += K3_THREADS / 32) {
            uint32_t vq = NAN2, vd = NAN2;
            if (j < Hw) {
                if (q0 + r < nq) vq = __ldg(q16 + (q0 + r) * (int64_t)Hw + j);
                if (d0 + r < nd) vd = __ldg(d16 + (d0 + r) * (int64_t)Hw + j);
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
                    asm("{ .reg .b32 e; set.eq.f16x2.f16x2 e, %1, %2; add.f16x2 %0, %0, e; }"
                        : "+r"(acc[a][b]) : "r"(q[a]), "r"(d[b]));
        }
        __syncthreads();
    }

    auto total = [](uint32_t h2) -> uint32_t {   // low half + high half, both exact small integers
        const __half2 v = *reinterpret_cast<const __half2*>(&h2);
        return (uint32_t)(__low2float(v) + __high2float(v));
    };
    const bool vec_ok = ((ld_out & 7) == 0) && ((((uintptr_t)cnt) & 15) == 0) && (d0 + tx * 8 + 8 <= nd);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t q = q0 + ty * 8 + a;
        if (q >= nq) continue;
        uint16_t* dst = cnt + q * ld_out + d0 + tx * 8;
        if (vec_ok) {
            uint4 o;
            o.x = total(acc[a][0]) | (total(acc[a][1]) << 16);
            o.y = total(acc[a][2]) | (total(acc[a][3]) << 16);
            o.z = total(acc[a][4]) | (total(acc[a][5]) << 16);
            o.w = total(acc[a][6]) | (total(acc[a][7]) << 16);
            *reinterpret_cast<uint4*>(dst) = o;
        } else {
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if (d0 + tx * 8 + b < nd) dst[b] = (uint16_t)total(acc[a][b]);
        }
    }
    if ((SYM && blockIdx.x > blockIdx.y) || MODE == 2) {
        uint16_t* __restrict__ mt = MODE == 2 ? cnt_t : cnt;
        const int64_t mld = MODE == 2 ? ld_t : ld_out;
        const bool tvec_ok = ((mld & 7) == 0) && ((((uintptr_t)mt) & 15) == 0) && (q0 + ty * 8 + 8 <= nq);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int64_t d = d0 + tx * 8 + b;
            if (d >= nd) continue;
            uint16_t* dst = mt + d * mld + q0 + ty * 8;
            if (tvec_ok) {
                uint4 o;
                o.x = total(acc[0][b]) | (total(acc[1][b]) << 16);
                o.y = total(acc[2][b]) | (total(acc[3][b]) << 16);
                o.z = total(acc[4][b]) | (total(acc[5][b]) << 16);
                o.w = total(acc[6][b]) | (total(acc[7][b]) << 16);
                *reinterpret_cast<uint4*>(dst) = o;
            } else {
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (q0 + ty * 8 + a < nq) dst[a] = (uint16_t)total(acc[a][b]);
            }
        }
    }
}

// =================================================================================================
// TMA-staged variant of the packed kernel (Blackwell bulk-tensor copies + mbarrier).
//
// The kernels above stage every 128 x 32 operand tile with LDG + transposed STS by all 256 threads and two barriers per
// step. Here the renamed signatures are ALSO kept transposed -- sig16T [Hw_pad][n_pad] uint32, row index contiguous
// (k3h_transpose_kernel) -- so that an operand tile [32 words][128 rows] is one 2-D box of a tensor map: one elected
// thread issues two cp.async.bulk.tensor.2d loads (query tile, database tile: 32 KB) per step into a double-buffered
// stage and arms an mbarrier with the byte count; the other threads only wait on the barrier's phase. Loads for step
// s + 1 are in flight while step s is compared, the compare warps issue no staging instructions at all, and one
// __syncthreads per step (stage reuse) remains instead of two. Row offsets must be multiples of 4 (a box starts on a
// 16-byte boundary; other offsets take the LDG kernel), rows past the end read as zero -- such rows and columns are never stored -- and the padding words
// (Hw .. Hw_pad) hold the never-equal NaN pattern.
// =================================================================================================
constexpr int K3T_STAGE_WORDS = 2 * K3_JB * K3_TILE;        // query tile + database tile, uint32

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "K3T_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra K3T_DONE;\n"
        "bra K3T_WAIT;\n"
        "K3T_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// sig16 [n][Hp] uint16 (row major, Hp even) -> sig16T [Hw_pad][n_pad] uint32; padding = NaN pattern
__global__ void __launch_bounds__(256)
k3h_transpose_kernel(const uint32_t* __restrict__ sig16w, int64_t n, int Hw, int Hw_pad, int64_t n_pad,
                     uint32_t* __restrict__ sig16T) {
    __shared__ uint32_t tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int j0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 8 x 32
    for (int a = ty; a < 32; a += 8) {
        const int64_t r = r0 + a;
        const int j = j0 + tx;
        tile[a][tx] = (r < n && j < Hw) ? __ldg(sig16w + r * (int64_t)Hw + j) : 0x7E007E00u;
    }
    __syncthreads();
    for (int a = ty; a < 32; a += 8) {
        const int j = j0 + a;
        const int64_t r = r0 + tx;
        if (j < Hw_pad && r < n_pad) sig16T[(int64_t)j * n_pad + r] = tile[tx][a];
    }
}

// One 128 x 128 tile (tile_x, tile_y) of the block [q_row0, +nq) x [d_row0, +nd). MODE as in k3h_collision_kernel.
template <int MODE>
__device__ __forceinline__ void k3t_tile(const CUtensorMap& tmap, int q_row0, int64_t nq, int d_row0, int64_t nd, int Hw_pad,
                                         uint16_t* __restrict__ cnt, int64_t ld_out, uint16_t* __restrict__ cnt_t, int64_t ld_t,
                                         int tile_x, int tile_y, uint32_t* stage_base, uint64_t* full_bar) {
    constexpr bool SYM = MODE == 1;
    if (SYM && tile_x < tile_y) return;
    const int64_t q0 = (int64_t)tile_y * K3_TILE;
    const int64_t d0 = (int64_t)tile_x * K3_TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int steps = Hw_pad / K3_JB;
    auto issue = [&](int s) {
        uint32_t* st = stage_base + (s & 1) * K3T_STAGE_WORDS;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic-proxy reads of this stage are done
        mbar_expect_tx(&full_bar[s & 1], K3T_STAGE_WORDS * 4);
        tma_load_2d(st, &tmap, (int)(q_row0 + q0), s * K3_JB, &full_bar[s & 1]);
        tma_load_2d(st + K3_JB * K3_TILE, &tmap, (int)(d_row0 + d0), s * K3_JB, &full_bar[s & 1]);
    };
    if (threadIdx.x == 0) issue(0);

    uint32_t acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0u;

    for (int s = 0; s < steps; ++s) {
        if (threadIdx.x == 0 && s + 1 < steps) issue(s + 1);        // its stage was released by the barrier ending step s - 1
        mbar_wait(&full_bar[s & 1], (uint32_t)((s >> 1) & 1));
        const uint32_t* Qs = stage_base + (s & 1) * K3T_STAGE_WORDS;
        const uint32_t* Ds = Qs + K3_JB * K3_TILE;
#pragma unroll 2
        for (int jj = 0; jj < K3_JB; ++jj) {
            const uint4 qa = *reinterpret_cast<const uint4*>(Qs + jj * K3_TILE + ty * 8);
            const uint4 qb = *reinterpret_cast<const uint4*>(Qs + jj * K3_TILE + ty * 8 + 4);
            const uint4 da = *reinterpret_cast<const uint4*>(Ds + jj * K3_TILE + tx * 8);
            const uint4 db = *reinterpret_cast<const uint4*>(Ds + jj * K3_TILE + tx * 8 + 4);
            const uint32_t q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const uint32_t d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b)
                    asm("{ .reg .b32 e; set.eq.f16x2.f16x2 e, %1, %2; add.f16x2 %0, %0, e; }"
                        : "+r"(acc[a][b]) : "r"(q[a]), "r"(d[b]));
        }
        __syncthreads();                                             // stage s & 1 may be refilled (by the issue of step s + 2)
    }

    auto total = [](uint32_t h2) -> uint32_t {
        const __half2 v = *reinterpret_cast<const __half2*>(&h2);
        return (uint32_t)(__low2float(v) + __high2float(v));
    };
    const bool vec_ok = ((ld_out & 7) == 0) && ((((uintptr_t)cnt) & 15) == 0) && (d0 + tx * 8 + 8 <= nd);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const int64_t q = q0 + ty * 8 + a;
        if (q >= nq) continue;
        uint16_t* dst = cnt + q * ld_out + d0 + tx * 8;
        if (vec_ok) {
            uint4 o;
            o.x = total(acc[a][0]) | (total(acc[a][1]) << 16);
            o.y = total(acc[a][2]) | (total(acc[a][3]) << 16);
            o.z = total(acc[a][4]) | (total(acc[a][5]) << 16);
            o.w = total(acc[a][6]) | (total(acc[a][7]) << 16);
            *reinterpret_cast<uint4*>(dst) = o;
        } else {
#pragma unroll
            for (int b = 0; b < 8; ++b)
                if (d0 + tx * 8 + b < nd) dst[b] = (uint16_t)total(acc[a][b]);
        }
    }
    if ((SYM && tile_x > tile_y) || MODE == 2) {
        uint16_t* __restrict__ mt = MODE == 2 ? cnt_t : cnt;
        const int64_t mld = MODE == 2 ? ld_t : ld_out;
        const bool tvec_ok = ((mld & 7) == 0) && ((((uintptr_t)mt) & 15) == 0) && (q0 + ty * 8 + 8 <= nq);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const int64_t d = d0 + tx * 8 + b;
            if (d >= nd) continue;
            uint16_t* dst = mt + d * mld + q0 + ty * 8;
            if (tvec_ok) {
                uint4 o;
                o.x = total(acc[0][b]) | (total(acc[1][b]) << 16);
                o.y = total(acc[2][b]) | (total(acc[3][b]) << 16);
                o.z = total(acc[4][b]) | (total(acc[5][b]) << 16);
                o.w = total(acc[6][b]) | (total(acc[7][b]) << 16);
                *reinterpret_cast<uint4*>(dst) = o;
            } else {
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    if (q0 + ty * 8 + a < nq) dst[a] = (uint16_t)total(acc[a][b]);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(K3_THREADS, 2)
k3t_collision_kernel(const __grid_constant__ CUtensorMap tmap, int q_row0, int64_t nq, int d_row0, int64_t nd, int Hw_pad,
                     uint16_t* __restrict__ cnt, int64_t ld_out, uint16_t* __restrict__ cnt_t, int64_t ld_t) {
    extern __shared__ __align__(1024) unsigned char k3t_smem[];
    __shared__ __align__(8) uint64_t full_bar[2];
    k3t_tile<MODE>(tmap, q_row0, nq, d_row0, nd, Hw_pad, cnt, ld_out, cnt_t, ld_t, (int)blockIdx.x, (int)blockIdx.y,
                   reinterpret_cast<uint32_t*>(k3t_smem), full_bar);
}

int k3t_padded_words(int H) { return ((k3h_padded_hashes(H) / 2) + K3_JB - 1) / K3_JB * K3_JB; }
int64_t k3t_padded_rows(int64_t n) { return (n + 3) & ~(int64_t)3; }
size_t k3t_transposed_elems(int64_t n, int H) { return (size_t)k3t_padded_words(H) * (size_t)k3t_padded_rows(n); }

// The tensor map over sig16T (driver entry point resolved through the runtime: no link against libcuda).
// out_map: 128 bytes (CUtensorMap). Returns 0 on success.
int k3t_make_tensor_map(const uint32_t* d_sig16T, int64_t n, int H, void* out_map) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            cudaGetLastError();
            return -1;
        }
        encode = (EncodeFn)fn;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)k3t_padded_rows(n), (cuuint64_t)k3t_padded_words(H)};
    const cuuint64_t strides[1] = {(cuuint64_t)k3t_padded_rows(n) * 4};
    const cuuint32_t box[2] = {(cuuint32_t)K3_TILE, (cuuint32_t)K3_JB};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode((CUtensorMap*)out_map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)d_sig16T, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -1;
}

int launch_transpose_signatures(const uint16_t* d_sig16, int64_t n, int H, uint32_t* d_sig16T, cudaStream_t stream) {
    if (n <= 0) return 0;
    const int Hw = k3h_padded_hashes(H) / 2, Hw_pad = k3t_padded_words(H);
    const int64_t n_pad = k3t_padded_rows(n);
    dim3 grid((unsigned)((n_pad + 31) / 32), (unsigned)(Hw_pad / 32));
    k3h_transpose_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(d_sig16), n, Hw, Hw_pad, n_pad, d_sig16T);
    return 1;
}

// Counts of database rows [q_row0, +nq) x [d_row0, +ndb) from the transposed renamed signatures behind `map`.
// Same rows twice (q_row0 == d_row0, nq == ndb, no transposed output): symmetric kernel.
int launch_collision_counts_tma(const void* map, int64_t q_row0, int64_t nq, int64_t d_row0, int64_t ndb, int H, uint16_t* d_cnt,
                                int64_t ld_out, uint16_t* d_cnt_t, int64_t ld_t, cudaStream_t stream) {
    if (nq <= 0 || ndb <= 0) return 0;
    const int Hw_pad = k3t_padded_words(H);
    dim3 grid((unsigned)((ndb + K3_TILE - 1) / K3_TILE), (unsigned)((nq + K3_TILE - 1) / K3_TILE));
    const size_t smem = (size_t)2 * K3T_STAGE_WORDS * 4 + 1024;
    const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(map);
    const bool sym = q_row0 == d_row0 && nq == ndb && !d_cnt_t;
    if (sym) {
        cudaFuncSetAttribute(k3t_collision_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k3t_collision_kernel<1><<<grid, K3_THREADS, smem, stream>>>(tm, (int)q_row0, nq, (int)d_row0, ndb, Hw_pad, d_cnt, ld_out, nullptr, 0);
    } else if (d_cnt_t) {
        cudaFuncSetAttribute(k3t_collision_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k3t_collision_kernel<2><<<grid, K3_THREADS, smem, stream>>>(tm, (int)q_row0, nq, (int)d_row0, ndb, Hw_pad, d_cnt, ld_out, d_cnt_t, ld_t);
    } else {
        cudaFuncSetAttribute(k3t_collision_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k3t_collision_kernel<0><<<grid, K3_THREADS, smem, stream>>>(tm, (int)q_row0, nq, (int)d_row0, ndb, Hw_pad, d_cnt, ld_out, nullptr, 0);
    }
    return 1;
}

bool k3h_applicable(int64_t ndb, int H) { return H <= 4094 && ndb <= K3H_MAX_ROWS && ndb > 0; }
int k3h_padded_hashes(int H) { return (H + 1) & ~1; }
size_t k3h_table_elems(int64_t ndb, int H) { return (size_t)H * (size_t)k3h_table_capacity(ndb); }

// d_table: k3h_table_elems uint32 (zeroed here), d_sig16: ndb * k3h_padded_hashes(H) uint16.
// max_bin_size > 0 with d_sizes (k3h_table_elems uint32, zeroed here): members of buckets with >= max_bin_size cells
// are blanked (rule 5) -- replaces K2 on this path.
int launch_rename_signatures(const uint32_t* d_sig_db, int64_t ndb, int H, uint32_t* d_table, uint16_t* d_sig16,
                             uint32_t* d_sizes, int64_t max_bin_size, cudaStream_t stream) {
    if (ndb <= 0) return 0;
    const int cap = k3h_table_capacity(ndb);
    const int Hp = k3h_padded_hashes(H);
    const bool prune = d_sizes && max_bin_size > 0;
    if (cudaMemsetAsync(d_table, 0, (size_t)H * cap * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    if (prune && cudaMemsetAsync(d_sizes, 0, (size_t)H * cap * sizeof(uint32_t), stream) != cudaSuccess) return -1;
    const int64_t total = ndb * (int64_t)Hp;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k3h_rename_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_sig_db, ndb, H, Hp, cap, d_table, prune ? d_sizes : nullptr, d_sig16);
    if (!prune) return 1;
    const uint32_t mb = (uint32_t)(max_bin_size > 0xFFFFFFFFll ? 0xFFFFFFFFll : max_bin_size);
    k3h_prune_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_sig16, ndb, H, Hp, cap, d_sizes, mb);
    return 2;
}

// q16 / d16: renamed signatures (rows of Hp uint16) of the query rows / of the database
int launch_collision_counts_packed(const uint16_t* d_q16, int64_t nq, const uint16_t* d_d16, int64_t ndb, int H,
                                   uint16_t* d_cnt, int64_t ld_out, cudaStream_t stream) {
    if (nq <= 0 || ndb <= 0) return 0;
    const int Hw = k3h_padded_hashes(H) / 2;
    dim3 grid((unsigned)((ndb + K3_TILE - 1) / K3_TILE), (unsigned)((nq + K3_TILE - 1) / K3_TILE));
    const uint32_t* q = reinterpret_cast<const uint32_t*>(d_q16);
    const uint32_t* d = reinterpret_cast<const uint32_t*>(d_d16);
    if (d_q16 == d_d16 && nq == ndb)
        k3h_collision_kernel<1><<<grid, K3_THREADS, 0, stream>>>(q, nq, d, ndb, Hw, d_cnt, ld_out, nullptr, 0);
    else
        k3h_collision_kernel<0><<<grid, K3_THREADS, 0, stream>>>(q, nq, d, ndb, Hw, d_cnt, ld_out, nullptr, 0);
    return 1;
}

// The same block with a transposed copy: d_cnt_t [ndb, ld_t] receives cnt_t[d][q] = cnt[q][d].
int launch_collision_counts_packed_t(const uint16_t* d_q16, int64_t nq, const uint16_t* d_d16, int64_t ndb, int H,
                                     uint16_t* d_cnt, int64_t ld_out, uint16_t* d_cnt_t, int64_t ld_t, cudaStream_t stream) {
    if (nq <= 0 || ndb <= 0) return 0;
    const int Hw = k3h_padded_hashes(H) / 2;
    dim3 grid((unsigned)((ndb + K3_TILE - 1) / K3_TILE), (unsigned)((nq + K3_TILE - 1) / K3_TILE));
    k3h_collision_kernel<2><<<grid, K3_THREADS, 0, stream>>>(reinterpret_cast<const uint32_t*>(d_q16), nq,
                                                             reinterpret_cast<const uint32_t*>(d_d16), ndb, Hw, d_cnt, ld_out,
                                                             d_cnt_t, ld_t);
    return 1;
}

}  // namespace schic
