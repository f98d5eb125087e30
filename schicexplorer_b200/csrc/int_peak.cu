// int_peak.cu -- INT32 throughput micro-benchmark (the roofline denominator of K1/K3).
//
// MEASURED_PEAKS.json holds only HBM and bf16 peaks; the MinHash kernels are bound by the integer
// issue rate, so bench.py measures that ceiling live with dependency-free chains of one SASS
// opcode each (LOP3 / SHF / VIMNMX on the ALU pipe, IMAD / IMAD.HI on the FMA pipe) and a 1:1
// ALU+FMA mix. 8 independent chains per thread, 8 resident CTAs of 256 threads per SM.
#include "common.cuh"

namespace schic {

enum PeakOp { OP_LOP3 = 0, OP_IMAD = 1, OP_MIX = 2, OP_SHF = 3, OP_IMADHI = 4, OP_MIN = 5, OP_MIN3 = 6, OP_ADD = 7, OP_WIDE = 8,
              OP_HSET2 = 9, OP_HADD2 = 10, OP_MIXH = 11, OP_MIXSETP = 12 };

template <int OP>
__device__ __forceinline__ void one_op(uint32_t& x, uint32_t y, uint32_t z, int chain) {
    if (OP == OP_LOP3 || (OP == OP_MIX && (chain & 1) == 0)) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    } else if (OP == OP_IMAD || OP == OP_MIX) {
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    } else if (OP == OP_SHF) {
        asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    } else if (OP == OP_IMADHI) {
        asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
    } else if (OP == OP_MIN) {
        asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
    } else if (OP == OP_MIN3) {
        asm volatile("{ .reg .u32 t; min.u32 t, %1, %2; min.u32 %0, %0, t; }" : "+r"(x) : "r"(y), "r"(z));
    } else if (OP == OP_HSET2 || (OP == OP_MIXH && (chain & 1) == 0)) {
        // packed fp16 compare producing 1.0h / 0.0h per half (K3 experiment: two 16-bit compares per instruction)
        asm volatile("set.eq.f16x2.f16x2 %0, %0, %1;" : "+r"(x) : "r"(y));
    } else if (OP == OP_HADD2 || OP == OP_MIXH) {
        asm volatile("add.f16x2 %0, %0, %1;" : "+r"(x) : "r"(z));
    } else if (OP == OP_MIXSETP) {
        // the current K3 inner pair: compare on the ALU pipe, predicated fp32 add on the FMA pipe
        asm volatile("{ .reg .pred p; .reg .f32 t; setp.eq.u32 p, %1, %2; mov.b32 t, %0; @p add.f32 t, t, 0f3F800000; mov.b32 %0, t; }"
                     : "+r"(x) : "r"(y), "r"(z + chain));
    } else if (OP == OP_WIDE) {
        asm volatile("{ .reg .u64 w; .reg .u32 lo; mul.wide.u32 w, %0, %1; mov.b64 {lo, %0}, w; }" : "+r"(x) : "r"(y));
    } else {
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
    }
}

template <int OP>
__global__ void __launch_bounds__(256)
peak_kernel(uint32_t* __restrict__ out, int iters, uint32_t y, uint32_t z, unsigned long long* __restrict__ clk) {
    uint32_t x[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) x[c] = threadIdx.x * 8u + c + blockIdx.x;
    unsigned long long t0 = 0, g0 = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        t0 = clock64();
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < 8; ++c) one_op<OP>(x[c], y, z, c);
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) s ^= x[c];
    if (s == 0x12345u) out[0] = s;   // keep the chains alive
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long g1;
        const unsigned long long t1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        clk[0] = t1 - t0;
        clk[1] = g1 - g0;
    }
}

template <int OP>
static double run_one(int iters, uint32_t* d_out, unsigned long long* d_clk, cudaStream_t stream, double* mhz) {
    const int grid = 148 * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    peak_kernel<OP><<<grid, block, 0, stream>>>(d_out, iters / 8 + 1, 0x9E3779B9u, 0x7F4A7C15u, d_clk);   // warm-up
    cudaEventRecord(e0, stream);
    peak_kernel<OP><<<grid, block, 0, stream>>>(d_out, iters, 0x9E3779B9u, 0x7F4A7C15u, d_clk);
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (mhz) {
        unsigned long long h[2] = {0, 1};
        cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
        *mhz = h[1] ? (double)h[0] / (double)h[1] * 1000.0 : 0.0;
    }
    const double ops = (double)grid * block * (double)iters * 32.0;   // 4 unroll x 8 chains
    return ops / (ms * 1e-3) / 1e12;
}

// out: [0] LOP3, [1] IMAD, [2] 1:1 LOP3+IMAD mix, [3] SM MHz during the mix, [4] SHF, [5] IMAD.HI,
//      [6] VIMNMX (2-input), [7] VIMNMX3 (counted as one op), [8] IADD, [9] IMAD.WIDE, [10] HSET2, [11] HADD2,
//      [12] 1:1 HSET2+HADD2, [13] ISETP + predicated FADD pairs (counted as one op per pair) -- all in T instr-lanes/s
int run_int32_peak(int iters, double* out, cudaStream_t stream) {
    uint32_t* d_out = nullptr;
    unsigned long long* d_clk = nullptr;
    if (cudaMalloc(&d_out, 64) != cudaSuccess) return -1;
    if (cudaMalloc(&d_clk, 64) != cudaSuccess) { cudaFree(d_out); return -1; }
    out[0] = run_one<OP_LOP3>(iters, d_out, d_clk, stream, nullptr);
    out[1] = run_one<OP_IMAD>(iters, d_out, d_clk, stream, nullptr);
    out[2] = run_one<OP_MIX>(iters, d_out, d_clk, stream, &out[3]);
    out[4] = run_one<OP_SHF>(iters, d_out, d_clk, stream, nullptr);
    out[5] = run_one<OP_IMADHI>(iters, d_out, d_clk, stream, nullptr);
    out[6] = run_one<OP_MIN>(iters, d_out, d_clk, stream, nullptr);
    out[7] = run_one<OP_MIN3>(iters, d_out, d_clk, stream, nullptr);
    out[8] = run_one<OP_ADD>(iters, d_out, d_clk, stream, nullptr);
    out[9] = run_one<OP_WIDE>(iters, d_out, d_clk, stream, nullptr);
    out[10] = run_one<OP_HSET2>(iters, d_out, d_clk, stream, nullptr);
    out[11] = run_one<OP_HADD2>(iters, d_out, d_clk, stream, nullptr);
    out[12] = run_one<OP_MIXH>(iters, d_out, d_clk, stream, nullptr);
    out[13] = run_one<OP_MIXSETP>(iters, d_out, d_clk, stream, nullptr);
    cudaFree(d_out);
    cudaFree(d_clk);
    return cudaGetLastError() == cudaSuccess ? 18 : -1;
}

}  // namespace schic
