// pca_dense.cu -- dense-graph PCA on the device (include/schic_pca.h).
//
// Replaces sklearn's PCA(n_components=min(shape)-1).fit_transform(graph.todense())[:, :dim] of
// scHicClusterMinHash.py:358-366 / scHicCluster.py:183-188. float64 throughout. Pipeline (one stream):
//   P1 densify / upload, column means, centring                                  (HBM-bound, one pass each)
//   P2 C = Xc^T Xc                     tiled DGEMM, upper tiles mirrored        (DFMA-bound)
//   P3 Householder tridiagonalisation  per step: one single-CTA kernel (reflector + w) and one pass over the
//      trailing matrix that applies the previous rank-2 update AND forms S v for the next step
//      (16 B per trailing element and step: HBM-bound, 16/3 n^3 bytes in total)
//   P4 the n_components largest eigenvalues: Sturm bisection, one thread per eigenvalue
//   P5 eigenvectors of the tridiagonal matrix: inverse iteration (pivoted LU, one thread per vector) with
//      classical Gram-Schmidt (twice) in eigenvalue order between sweeps
//   P6 back-transformation (one CTA per vector walks the reflectors), sklearn's sign convention, out = Xc V.
// The numpy restatement of exactly these steps is oracle/pca_np.py (pinned to scikit-learn).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/schic_mh.h"
#include "../../include/schic_pca.h"
#include "common.cuh"

namespace {

#define PCA_CU(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            cudaGetLastError();                                                                        \
            return schic::set_error(e__ == cudaErrorMemoryAllocation ? SCHIC_ERR_OOM : SCHIC_ERR_CUDA, \
                                    std::string(#expr) + ": " + cudaGetErrorString(e__));              \
        }                                                                                              \
    } while (0)

template <typename T>
struct Buf {
    T* p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA (blockDim.x multiple of 32, <= 1024); every thread gets the result. red: 33 doubles of shared memory.
__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();                       // red may still be read from a previous call
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// ---- P1 ---------------------------------------------------------------------------------------------------
__global__ void pca_scatter_csr_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                       const float* __restrict__ data, int64_t n_rows, int64_t n_cols,
                                       double* __restrict__ X) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int lane = threadIdx.x & 31;
    for (int64_t e = indptr[row] + lane; e < indptr[row + 1]; e += 32) {
        const int32_t c = indices[e];
        const float v = data[e];
        if (c >= 0 && c < n_cols) X[row * n_cols + c] = isfinite(v) ? (double)v : 0.0;
    }
}

__global__ void pca_scrub_kernel(double* __restrict__ X, int64_t total) {      // dense input: nan / inf -> 0
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        if (!isfinite(X[i])) X[i] = 0.0;
}

__global__ void pca_col_sum_kernel(const double* __restrict__ X, int64_t n_rows, int64_t n_cols, double* __restrict__ mu) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const int64_t per = (n_rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per, r1 = min(n_rows, r0 + per);
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) s += X[r * n_cols + c];
    atomicAdd(&mu[c], s);
}

__global__ void pca_center_kernel(double* __restrict__ X, int64_t n_rows, int64_t n_cols, const double* __restrict__ mu) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const double m = mu[c] / (double)n_rows;
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) X[r * n_cols + c] -= m;
}

// ---- P2 / P6: tiled DGEMM --------------------------------------------------------------------------------------
// out[i][j] = sum_k A(i,k) * B(j,k), i < M, j < N, k < K.
//   KMAJOR = true : operands stored [k][index]  (A(i,k) = A[k*lda + i]) -- the covariance Xc^T Xc, with SYM: only
//                   tiles bj >= bi are computed and stored to both halves;
//   KMAJOR = false: operands stored [index][k]  (A(i,k) = A[i*lda + k]) -- the projection Xc V^T-layout.
constexpr int GT = 64, GK = 16;
template <bool KMAJOR, bool SYM>
__global__ void __launch_bounds__(256)
pca_gemm_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb, int64_t M, int64_t N,
                int64_t K, double* __restrict__ out, int64_t ldo) {
    if (SYM && blockIdx.x < blockIdx.y) return;
    __shared__ double As[GK][GT + 1], Bs[GK][GT + 1];
    const int64_t i0 = (int64_t)blockIdx.y * GT, j0 = (int64_t)blockIdx.x * GT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4] = {};
    for (int64_t k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = threadIdx.x + 256 * u;
            if (KMAJOR) {
                const int kk = e >> 6, ii = e & 63;
                const int64_t k = k0 + kk;
                As[kk][ii] = (k < K && i0 + ii < M) ? A[k * lda + i0 + ii] : 0.0;
                Bs[kk][ii] = (k < K && j0 + ii < N) ? B[k * ldb + j0 + ii] : 0.0;
            } else {
                const int ii = e >> 4, kk = e & 15;
                const int64_t k = k0 + kk;
                As[kk][ii] = (k < K && i0 + ii < M) ? A[(i0 + ii) * lda + k] : 0.0;
                Bs[kk][ii] = (k < K && j0 + ii < N) ? B[(j0 + ii) * ldb + k] : 0.0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[kk][ty + 16 * u]; b[u] = Bs[kk][tx + 16 * u]; }   // conflict-free: consecutive lanes, consecutive doubles
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int64_t i = i0 + ty + 16 * u, j = j0 + tx + 16 * v;
            if (i < M && j < N) {
                out[i * ldo + j] = acc[u][v];
                if (SYM && blockIdx.x != blockIdx.y) out[j * ldo + i] = acc[u][v];
            }
        }
}

// ---- P3: tridiagonalisation ----------------------------------------------------------------------------------
// Step j (-1 .. n-2), single CTA:
//   w_j = tau_j S v_j - (tau_j^2 (v_j^T S v_j) / 2) v_j          (praw = S v_j comes from the previous pass; j = -1: 0)
//   x   = row r = j+1 of S with the pending rank-2 update (v_j, w_j) applied;  d[r] = x[r]
//   reflector r from x[r+1:]: beta, tau_r, v_r (v_r[r+1] = 1); e[r] = beta; v_r is also kept in S[r][r+2:].
// LOWER: the passes keep only the lower triangle up to date (pca_tridiag_lower_pass_kernel), so row r is read as
// column r below the diagonal, and praw -- accumulated there with atomics -- is zeroed here once consumed.
template <bool LOWER>
__global__ void __launch_bounds__(1024)
pca_tridiag_step_kernel(double* __restrict__ S, int64_t n, int64_t j, double* __restrict__ praw,
                        const double* __restrict__ vcur, double* __restrict__ w, double* __restrict__ vnext,
                        double* __restrict__ d, double* __restrict__ e, double* __restrict__ tau) {
    __shared__ double red[33];
    const int64_t r = j + 1;
    if (j >= 0) {
        const double tj = tau[j];
        double part = 0.0;
        for (int64_t i = r + threadIdx.x; i < n; i += blockDim.x) part += praw[i] * vcur[i];
        const double alpha = block_sum(part, red);
        const double coef = 0.5 * tj * tj * alpha;
        for (int64_t i = r + threadIdx.x; i < n; i += blockDim.x) {
            w[i] = tj * praw[i] - coef * vcur[i];
            if (LOWER) praw[i] = 0.0;
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) w[i] = 0.0;
    }
    __syncthreads();
    const double vr = j >= 0 ? vcur[r] : 0.0, wr = w[r];
    double* __restrict__ row = S + r * n;
    double part = 0.0;
    for (int64_t k = r + threadIdx.x; k < n; k += blockDim.x) {
        const double x = (LOWER ? S[k * n + r] : row[k]) - vr * w[k] - wr * (j >= 0 ? vcur[k] : 0.0);
        vnext[k] = x;
        if (k >= r + 2) part += x * x;
    }
    const double xnorm2 = block_sum(part, red);      // (its barriers also publish vnext)
    if (threadIdx.x == 0) d[r] = vnext[r];
    if (r == n - 1) return;
    const double alpha = vnext[r + 1];
    __syncthreads();                                 // everyone has read alpha before vnext[r+1] is overwritten
    if (xnorm2 == 0.0) {
        if (threadIdx.x == 0) { tau[r] = 0.0; e[r] = alpha; vnext[r + 1] = 1.0; }
        for (int64_t k = r + 2 + threadIdx.x; k < n; k += blockDim.x) { vnext[k] = 0.0; row[k] = 0.0; }
        return;
    }
    const double beta = -copysign(hypot(alpha, sqrt(xnorm2)), alpha);
    const double scale = 1.0 / (alpha - beta);
    if (threadIdx.x == 0) { tau[r] = (beta - alpha) / beta; e[r] = beta; vnext[r + 1] = 1.0; }
    for (int64_t k = r + 2 + threadIdx.x; k < n; k += blockDim.x) {
        const double v = vnext[k] * scale;
        vnext[k] = v;
        row[k] = v;
    }
}

// The pass over the trailing matrix of step j: rows / columns >= t0 = j + 2. One warp per row:
//   S[i][k] -= v_j[i] w_j[k] + w_j[i] v_j[k];   praw[i] = sum_k S_new[i][k] v_{j+1}[k]
__global__ void __launch_bounds__(256)
pca_tridiag_pass_kernel(double* __restrict__ S, int64_t n, int64_t t0, const double* __restrict__ vcur,
                        const double* __restrict__ w, const double* __restrict__ vnext, double* __restrict__ praw) {
    const int lane = threadIdx.x & 31;
    const int64_t i = t0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const double vi = vcur[i], wi = w[i];
    double* __restrict__ row = S + i * n;
    double acc = 0.0;
    for (int64_t k = t0 + lane; k < n; k += 32) {
        const double s = row[k] - vi * w[k] - wi * vcur[k];
        row[k] = s;
        acc = fma(s, vnext[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) praw[i] = acc;
}

// The same pass on the lower triangle only (half the traffic): rows i >= t0, columns t0 <= k <= i.
//   S[i][k] -= v_j[i] w_j[k] + w_j[i] v_j[k];   praw[i] += S_new[i][k] v_{j+1}[k];   praw[k] += S_new[i][k] v_{j+1}[i] (k < i)
// Grid (band of 32 rows, chunk of 8 column tiles); a warp owns one 32 x 32 tile, a lane one of its columns: the column
// sum stays in a register over the 32 rows, the 32 row sums are transposed across the warp at the end (31 shuffles), so
// a tile costs 2 x 32 atomics. Short CTAs keep the per-step critical path at a few load round trips.
__device__ __forceinline__ double tile_row_sums(double (&racc)[32], int lane) {
    // after round h (16, 8, 4, 2, 1) a lane keeps the accumulators whose index agrees with its lane bits so far;
    // at the end lane l holds the warp-wide sum of row l
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const bool upper = (lane & h) != 0;
#pragma unroll
        for (int u = 0; u < h; ++u) {
            const double send = upper ? racc[u] : racc[u + h];
            const double keep = upper ? racc[u + h] : racc[u];
            racc[u] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return racc[0];
}

__global__ void __launch_bounds__(256, 2)
pca_tridiag_lower_pass_kernel(double* __restrict__ S, int64_t n, int64_t t0, const double* __restrict__ vcur,
                              const double* __restrict__ w, const double* __restrict__ vnext, double* __restrict__ praw) {
    __shared__ double vs[32], ws[32], vns[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = t0 + 32 * (int64_t)blockIdx.x;
    const int rows = (int)min((int64_t)32, n - i0);
    const int64_t kfirst = t0 & ~(int64_t)31;                 // tiles aligned to 256 B
    const int64_t k0 = kfirst + 32 * ((int64_t)blockIdx.y * 8 + warp);
    if (kfirst + 32 * (int64_t)blockIdx.y * 8 > i0 + rows - 1) return;      // whole CTA right of the diagonal block
    if (threadIdx.x < 32) {
        const int64_t i = i0 + threadIdx.x;
        vs[threadIdx.x] = i < n ? vcur[i] : 0.0;
        ws[threadIdx.x] = i < n ? w[i] : 0.0;
        vns[threadIdx.x] = i < n ? vnext[i] : 0.0;
    }
    __syncthreads();
    if (k0 > i0 + rows - 1) return;
    const int64_t k = k0 + lane;
    const bool kin = k >= t0 && k < n;
    const double wk = kin ? w[k] : 0.0, vk = kin ? vcur[k] : 0.0, vnk = kin ? vnext[k] : 0.0;
    const bool inner = k0 + 31 < i0 && k0 >= t0;              // tile entirely left of the band's diagonal block
    double racc[32];
    double cacc = 0.0;
    double* __restrict__ gp = S + i0 * n + k;                // walks down the band's rows
    // rows in groups of 8: all loads of a group are issued before its first store (the compiler keeps loads behind
    // earlier stores through the same base pointer, which would leave one load in flight per warp)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        double x[8];
        bool ok[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = 8 * g + u;
            ok[u] = r < rows && (inner || (kin && k <= i0 + r));
            x[u] = ok[u] ? gp[u * n] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = 8 * g + u;
            double ra = 0.0;
            if (ok[u]) {
                const double sv = x[u] - vs[r] * wk - ws[r] * vk;
                gp[u * n] = sv;
                ra = sv * vnk;
                if (inner || k < i0 + r) cacc = fma(sv, vns[r], cacc);
            }
            racc[r] = ra;
        }
        gp += 8 * n;
    }
    if (kin && cacc != 0.0) atomicAdd(&praw[k], cacc);
    const double rsum = tile_row_sums(racc, lane);
    if (lane < rows && rsum != 0.0) atomicAdd(&praw[i0 + lane], rsum);
}

// ---- P4: Sturm bisection ----------------------------------------------------------------------------------------
// theta[t] = (t+1)-th largest eigenvalue of the symmetric tridiagonal (d, e). One thread per eigenvalue.
__global__ void pca_bisect_kernel(const double* __restrict__ d, const double* __restrict__ e, int64_t n, int m,
                                  double* __restrict__ theta) {
    __shared__ double smax[32];
    double bnd = 0.0;                              // Gershgorin bound of the spectrum (every CTA computes it)
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double a = fabs(d[i]) + (i > 0 ? fabs(e[i - 1]) : 0.0) + (i + 1 < n ? fabs(e[i]) : 0.0);
        bnd = fmax(bnd, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bnd = fmax(bnd, __shfl_xor_sync(0xffffffffu, bnd, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = bnd;
    __syncthreads();
    bnd = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) bnd = fmax(bnd, smax[i]);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int64_t target = n - 1 - t;              // ascending index of this eigenvalue
    double lo = -bnd, hi = bnd;
    const double tiny = 1e-300;
    for (int it = 0; it < 110; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        int64_t cnt = 0;                           // number of eigenvalues < mid
        double q = d[0] - mid;
        cnt += q < 0.0;
        for (int64_t i = 1; i < n; ++i) {
            if (fabs(q) < tiny) q = -tiny;
            const double ei = e[i - 1];
            q = d[i] - mid - ei * ei / q;
            cnt += q < 0.0;
        }
        if (cnt <= target) lo = mid; else hi = mid;
    }
    theta[t] = 0.5 * (lo + hi);
}

// ---- P5: inverse iteration -----------------------------------------------------------------------------------------
// One thread per vector: LU of (T - theta I) with partial pivoting (dgtsv layout), solve with the current
// column of X as the right-hand side, normalise. X is [n][m] (row = coordinate), ws = 3 arrays [n][m].
__global__ void pca_invit_kernel(const double* __restrict__ d, const double* __restrict__ e, int64_t n, int m,
                                 const double* __restrict__ theta, double* __restrict__ X, double* __restrict__ ws,
                                 double pivmin) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const double th = theta[t];
    double* __restrict__ DD = ws;                       // pivots
    double* __restrict__ DU = ws + n * (int64_t)m;      // first super-diagonal of U
    double* __restrict__ DU2 = ws + 2 * n * (int64_t)m; // second super-diagonal of U
    auto at = [&](int64_t i) { return i * (int64_t)m + t; };
    double dd = d[0] - th;                              // running diagonal entry i
    double du = n > 1 ? e[0] : 0.0;                     // running super-diagonal entry i
    double xi = X[at(0)];
    for (int64_t i = 0; i + 1 < n; ++i) {
        const double dl = e[i];
        const double dnext = d[i + 1] - th;
        const double unext = i + 2 < n ? e[i + 1] : 0.0;
        double xn = X[at(i + 1)];
        if (fabs(dd) >= fabs(dl)) {                     // no row interchange
            if (dd == 0.0) dd = pivmin;
            const double f = dl / dd;
            DD[at(i)] = dd; DU[at(i)] = du; DU2[at(i)] = 0.0;
            X[at(i)] = xi;
            xi = xn - f * xi;
            dd = dnext - f * du;
            du = unext;
        } else {                                        // interchange rows i and i+1
            const double f = dd / dl;
            DD[at(i)] = dl; DU[at(i)] = dnext; DU2[at(i)] = unext;
            X[at(i)] = xn;
            xi = xi - f * xn;
            dd = du - f * dnext;
            du = -f * unext;
        }
    }
    if (dd == 0.0) dd = pivmin;
    // back substitution
    double x1 = xi / dd;                                // x[n-1]
    double ss = x1 * x1;
    X[at(n - 1)] = x1;
    double x2 = 0.0;
    for (int64_t i = n - 2; i >= 0; --i) {
        const double x0 = (X[at(i)] - DU[at(i)] * x1 - DU2[at(i)] * x2) / DD[at(i)];
        X[at(i)] = x0;
        ss += x0 * x0;
        x2 = x1; x1 = x0;
    }
    // normalise (guard against overflow of ss: rescale by the largest entry first when needed)
    double inv = 1.0 / sqrt(ss);
    if (!isfinite(inv) || inv == 0.0) {
        double mx = 0.0;
        for (int64_t i = 0; i < n; ++i) mx = fmax(mx, fabs(X[at(i)]));
        double s2 = 0.0;
        for (int64_t i = 0; i < n; ++i) { const double v = X[at(i)] / mx; s2 += v * v; }
        inv = 1.0 / (mx * sqrt(s2));
    }
    for (int64_t i = 0; i < n; ++i) X[at(i)] *= inv;
}

__global__ void pca_init_vectors_kernel(double* __restrict__ X, int64_t n, int m) {   // deterministic start vectors
    const int64_t total = n * (int64_t)m;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = (uint64_t)i * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;      // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        X[i] = ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 0.5;
    }
}

// Gram-Schmidt of column col against columns [0, col): c[p] += sum over this CTA's rows of X[r][p] X[r][col]
__global__ void __launch_bounds__(256)
pca_gs_dots_kernel(const double* __restrict__ X, int64_t n, int m, int col, double* __restrict__ c) {
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(n, r0 + per);
    for (int p = threadIdx.x; p < col; p += blockDim.x) {
        double s = 0.0;
        for (int64_t r = r0; r < r1; ++r) s = fma(X[r * m + p], X[r * m + col], s);
        atomicAdd(&c[p], s);
    }
}

// X[r][col] -= sum_p X[r][p] c[p]; nrm2 += the new column's squares. One warp per row.
__global__ void __launch_bounds__(256)
pca_gs_apply_kernel(double* __restrict__ X, int64_t n, int m, int col, const double* __restrict__ c, double* __restrict__ nrm2) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    double sq = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps) {
        double s = 0.0;
        for (int p = lane; p < col; p += 32) s = fma(X[r * m + p], c[p], s);
        s = warp_sum(s);
        const double x = X[r * m + col] - s;
        if (lane == 0) { X[r * m + col] = x; sq += x * x; }
    }
    if (lane == 0) atomicAdd(nrm2, sq);
}

__global__ void pca_col_scale_kernel(double* __restrict__ X, int64_t n, int m, int col, const double* __restrict__ nrm2) {
    const double inv = *nrm2 > 0.0 ? 1.0 / sqrt(*nrm2) : 0.0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        X[r * m + col] *= inv;
}

// ---- P6 --------------------------------------------------------------------------------------------------------------
__global__ void pca_transpose_kernel(const double* __restrict__ X, int64_t n, int m, double* __restrict__ Z) {   // [n][m] -> [m][n]
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int64_t r = r0 + y; const int c = c0 + threadIdx.x;
        tile[y][threadIdx.x] = (r < n && c < m) ? X[r * m + c] : 0.0;
    }
    __syncthreads();
    for (int y = threadIdx.y; y < 32; y += blockDim.y) {
        const int c = c0 + y; const int64_t r = r0 + threadIdx.x;
        if (c < m && r < n) Z[(int64_t)c * n + r] = tile[threadIdx.x][y];
    }
}

// z <- H_0 H_1 ... H_{n-2} z for the CTA's vector (reflector r: v[r+1] = 1, v[k] = S[r][k] for k >= r+2), then the
// sign convention: the entry of largest magnitude (first one on ties) becomes positive.
__global__ void __launch_bounds__(256)
pca_back_transform_kernel(const double* __restrict__ S, const double* __restrict__ tau, int64_t n, double* __restrict__ Z,
                          int use_smem) {
    extern __shared__ double zs[];
    __shared__ double red[33];
    __shared__ int64_t s_arg;
    double* __restrict__ zg = Z + (int64_t)blockIdx.x * n;
    double* z = use_smem ? zs : zg;
    if (use_smem) {
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) zs[i] = zg[i];
        __syncthreads();
    }
    for (int64_t r = n - 2; r >= 0; --r) {
        const double t = tau[r];
        if (t == 0.0) continue;                            // uniform over the CTA
        const double* __restrict__ v = S + r * n;
        double part = threadIdx.x == 0 ? z[r + 1] : 0.0;
        for (int64_t k = r + 2 + threadIdx.x; k < n; k += blockDim.x) part = fma(v[k], z[k], part);
        const double dot = block_sum(part, red) * t;
        if (threadIdx.x == 0) z[r + 1] -= dot;
        for (int64_t k = r + 2 + threadIdx.x; k < n; k += blockDim.x) z[k] = fma(-dot, v[k], z[k]);
        // (the next iteration's block_sum starts with a barrier: the updates above are ordered before its reads
        //  of other threads' entries only through that barrier, so issue one here)
        __syncthreads();
    }
    // argmax |z| (first occurrence)
    double best = -1.0; int64_t arg = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double a = fabs(z[i]);
        if (a > best) { best = a; arg = i; }
    }
    __shared__ double sb[256];
    __shared__ int64_t sa[256];
    sb[threadIdx.x] = best; sa[threadIdx.x] = arg;
    __syncthreads();
    if (threadIdx.x == 0) {
        double b = -1.0; int64_t a = 0;
        for (int i = 0; i < (int)blockDim.x; ++i)
            if (sb[i] > b || (sb[i] == b && sa[i] < a)) { b = sb[i]; a = sa[i]; }
        s_arg = a;
    }
    __syncthreads();
    const double sign = z[s_arg] < 0.0 ? -1.0 : 1.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) zg[i] = sign * z[i];
}

__global__ void pca_sqrt_kernel(const double* __restrict__ theta, int m, double* __restrict__ sv) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) sv[t] = sqrt(fmax(theta[t], 0.0));
}

thread_local float g_stage_ms[7] = {0, 0, 0, 0, 0, 0, 0};

struct Events {
    cudaEvent_t ev[7] = {};
    ~Events() { for (auto& e : ev) if (e) cudaEventDestroy(e); }
};

// The whole pipeline on a device-resident, row-major X [n_rows, n_cols] (centred in place).
int pca_device(double* X, int64_t n_rows, int64_t n_cols, int m, cudaStream_t s, Events& ev, double* out_host,
               double* sv_host, double* comp_host, double* mean_host) {
    const int64_t n = n_cols;
    Buf<double> mu, C, praw, va, vb, w, d, e, tau, theta, Xv, ws, gsc, Z, outd, svd;
    PCA_CU(mu.alloc(n)); PCA_CU(C.alloc((size_t)n * n)); PCA_CU(praw.alloc(n)); PCA_CU(va.alloc(n)); PCA_CU(vb.alloc(n));
    PCA_CU(w.alloc(n)); PCA_CU(d.alloc(n)); PCA_CU(e.alloc(n)); PCA_CU(tau.alloc(n)); PCA_CU(theta.alloc(m));
    PCA_CU(Xv.alloc((size_t)n * m)); PCA_CU(ws.alloc((size_t)3 * n * m)); PCA_CU(gsc.alloc(m + 1));
    PCA_CU(Z.alloc((size_t)n * m)); PCA_CU(outd.alloc((size_t)n_rows * m)); PCA_CU(svd.alloc(m));

    // P1: centre
    PCA_CU(cudaMemsetAsync(mu.p, 0, n * sizeof(double), s));
    {
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)std::min<int64_t>(64, std::max<int64_t>(1, n_rows / 64)));
        pca_col_sum_kernel<<<grid, 256, 0, s>>>(X, n_rows, n, mu.p);
        pca_center_kernel<<<grid, 256, 0, s>>>(X, n_rows, n, mu.p);
    }
    PCA_CU(cudaEventRecord(ev.ev[1], s));
    // P2: covariance (unnormalised): C = Xc^T Xc
    {
        dim3 grid((unsigned)((n + GT - 1) / GT), (unsigned)((n + GT - 1) / GT));
        pca_gemm_kernel<true, true><<<grid, 256, 0, s>>>(X, n, X, n, n, n, n_rows, C.p, n);
    }
    PCA_CU(cudaGetLastError());
    PCA_CU(cudaEventRecord(ev.ev[2], s));
    // P3: tridiagonalisation
    PCA_CU(cudaMemsetAsync(va.p, 0, n * sizeof(double), s));
    PCA_CU(cudaMemsetAsync(vb.p, 0, n * sizeof(double), s));
    PCA_CU(cudaMemsetAsync(praw.p, 0, n * sizeof(double), s));
    PCA_CU(cudaMemsetAsync(tau.p, 0, n * sizeof(double), s));
    PCA_CU(cudaMemsetAsync(e.p, 0, n * sizeof(double), s));
    {
        double* vcur = va.p; double* vnext = vb.p;
        const char* mode = getenv("SCHIC_PCA_TRIDIAG");          // "full": the full-square pass (tests compare the two)
        const bool lower = !(mode && strcmp(mode, "full") == 0);
        for (int64_t j = -1; j <= n - 2; ++j) {
            if (lower) pca_tridiag_step_kernel<true><<<1, 1024, 0, s>>>(C.p, n, j, praw.p, vcur, w.p, vnext, d.p, e.p, tau.p);
            else pca_tridiag_step_kernel<false><<<1, 1024, 0, s>>>(C.p, n, j, praw.p, vcur, w.p, vnext, d.p, e.p, tau.p);
            const int64_t t0 = j + 2;
            if (t0 < n) {
                const int64_t rows = n - t0;
                if (lower) {
                    const int64_t nb = (rows + 31) / 32;       // bands; band b has at most b + 2 column tiles
                    dim3 grid((unsigned)nb, (unsigned)((nb + 1 + 7) / 8));
                    pca_tridiag_lower_pass_kernel<<<grid, 256, 0, s>>>(C.p, n, t0, vcur, w.p, vnext, praw.p);
                }
                else
                    pca_tridiag_pass_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(C.p, n, t0, vcur, w.p, vnext, praw.p);
            }
            std::swap(vcur, vnext);
        }
    }
    PCA_CU(cudaGetLastError());
    PCA_CU(cudaEventRecord(ev.ev[3], s));
    // P4: eigenvalues
    pca_bisect_kernel<<<(m + 127) / 128, 128, 0, s>>>(d.p, e.p, n, m, theta.p);
    PCA_CU(cudaGetLastError());
    PCA_CU(cudaEventRecord(ev.ev[4], s));
    // P5: eigenvectors of T
    {
        // pivmin: eps * ||T||_1, read back (two small vectors)
        std::vector<double> hd(n), he(n);
        PCA_CU(cudaMemcpyAsync(hd.data(), d.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
        PCA_CU(cudaMemcpyAsync(he.data(), e.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
        PCA_CU(cudaStreamSynchronize(s));
        double tn = 0.0;
        for (int64_t i = 0; i < n; ++i)
            tn = std::max(tn, std::fabs(hd[i]) + (i > 0 ? std::fabs(he[i - 1]) : 0.0) + (i + 1 < n ? std::fabs(he[i]) : 0.0));
        const double pivmin = (tn > 0.0 ? tn : 1.0) * 2.220446049250313e-16;   // tn == 0: the all-zero matrix (any basis is right)
        pca_init_vectors_kernel<<<148 * 4, 256, 0, s>>>(Xv.p, n, m);
        const int gs_blocks = (int)std::min<int64_t>(148 * 2, std::max<int64_t>(1, n / 32));
        for (int sweep = 0; sweep < 3; ++sweep) {
            pca_invit_kernel<<<(m + 63) / 64, 64, 0, s>>>(d.p, e.p, n, m, theta.p, Xv.p, ws.p, pivmin);
            for (int col = 0; col < m; ++col) {
                for (int pass = 0; pass < 2 && col > 0; ++pass) {
                    PCA_CU(cudaMemsetAsync(gsc.p, 0, (m + 1) * sizeof(double), s));
                    pca_gs_dots_kernel<<<gs_blocks, 256, 0, s>>>(Xv.p, n, m, col, gsc.p);
                    pca_gs_apply_kernel<<<gs_blocks, 256, 0, s>>>(Xv.p, n, m, col, gsc.p, gsc.p + m);
                }
                if (col == 0) {
                    PCA_CU(cudaMemsetAsync(gsc.p, 0, (m + 1) * sizeof(double), s));
                    pca_gs_apply_kernel<<<gs_blocks, 256, 0, s>>>(Xv.p, n, m, 0, gsc.p, gsc.p + m);   // norm only
                }
                pca_col_scale_kernel<<<gs_blocks, 256, 0, s>>>(Xv.p, n, m, col, gsc.p + m);
            }
        }
    }
    PCA_CU(cudaGetLastError());
    PCA_CU(cudaEventRecord(ev.ev[5], s));
    // P6: back-transform, sign, projection
    {
        dim3 tb(32, 8), tg((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32));
        pca_transpose_kernel<<<tg, tb, 0, s>>>(Xv.p, n, m, Z.p);
        const size_t zbytes = (size_t)n * sizeof(double);
        const int use_smem = zbytes <= 200 * 1024;
        if (use_smem) PCA_CU(cudaFuncSetAttribute(pca_back_transform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zbytes));
        pca_back_transform_kernel<<<m, 256, use_smem ? zbytes : 0, s>>>(C.p, tau.p, n, Z.p, use_smem);
        dim3 grid((unsigned)((m + GT - 1) / GT), (unsigned)((n_rows + GT - 1) / GT));
        pca_gemm_kernel<false, false><<<grid, 256, 0, s>>>(X, n, Z.p, n, n_rows, m, n, outd.p, m);
        pca_sqrt_kernel<<<(m + 127) / 128, 128, 0, s>>>(theta.p, m, svd.p);
    }
    PCA_CU(cudaGetLastError());
    PCA_CU(cudaMemcpyAsync(out_host, outd.p, (size_t)n_rows * m * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (sv_host) PCA_CU(cudaMemcpyAsync(sv_host, svd.p, m * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (comp_host) PCA_CU(cudaMemcpyAsync(comp_host, Z.p, (size_t)n * m * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (mean_host) PCA_CU(cudaMemcpyAsync(mean_host, mu.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));   // column sums
    PCA_CU(cudaEventRecord(ev.ev[6], s));
    PCA_CU(cudaStreamSynchronize(s));
    if (mean_host)
        for (int64_t c = 0; c < n; ++c) mean_host[c] /= (double)n_rows;
    return SCHIC_OK;
}

int pca_entry(const double* X_host, int64_t n_rows, int64_t n_cols, const int64_t* indptr, const int32_t* indices,
              const float* data, int32_t n_components, int32_t device, double* out, double* sv, double* comp,
              double* mean) {
    if (n_rows < 1 || n_cols < 1 || !out) return schic::set_error(SCHIC_ERR_INVALID_ARG, "bad PCA arguments");
    if (!X_host && !indptr) return schic::set_error(SCHIC_ERR_INVALID_ARG, "no input matrix");
    if (n_components < 1 || n_components > std::min(n_rows, n_cols))
        return schic::set_error(SCHIC_ERR_INVALID_ARG, "n_components must be in [1, min(n_rows, n_cols)]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return schic::set_error(SCHIC_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    }
    if (device >= 0) PCA_CU(cudaSetDevice(device));
    struct Stream {
        cudaStream_t s = nullptr;
        ~Stream() { if (s) cudaStreamDestroy(s); }
    } stream;
    PCA_CU(cudaStreamCreateWithFlags(&stream.s, cudaStreamNonBlocking));
    cudaStream_t s = stream.s;
    Events ev;
    for (auto& e : ev.ev) PCA_CU(cudaEventCreate(&e));
    int rc = SCHIC_OK;
    {
        Buf<double> X;
        Buf<int64_t> d_indptr; Buf<int32_t> d_indices; Buf<float> d_data;
        auto body = [&]() -> int {
            const size_t total = (size_t)n_rows * n_cols;
            PCA_CU(X.alloc(total));
            PCA_CU(cudaEventRecord(ev.ev[0], s));
            if (X_host) {
                PCA_CU(cudaMemcpyAsync(X.p, X_host, total * sizeof(double), cudaMemcpyHostToDevice, s));
                pca_scrub_kernel<<<148 * 8, 256, 0, s>>>(X.p, (int64_t)total);
            } else {
                const int64_t nnz = indptr[n_rows] - indptr[0];
                if (indptr[0] != 0 || nnz < 0 || (nnz > 0 && (!indices || !data)))
                    return schic::set_error(SCHIC_ERR_INVALID_ARG, "bad CSR arguments");
                PCA_CU(d_indptr.alloc(n_rows + 1)); PCA_CU(d_indices.alloc(nnz)); PCA_CU(d_data.alloc(nnz));
                PCA_CU(cudaMemcpyAsync(d_indptr.p, indptr, (n_rows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
                if (nnz > 0) {
                    PCA_CU(cudaMemcpyAsync(d_indices.p, indices, nnz * sizeof(int32_t), cudaMemcpyHostToDevice, s));
                    PCA_CU(cudaMemcpyAsync(d_data.p, data, nnz * sizeof(float), cudaMemcpyHostToDevice, s));
                }
                PCA_CU(cudaMemsetAsync(X.p, 0, total * sizeof(double), s));
                pca_scatter_csr_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, s>>>(d_indptr.p, d_indices.p, d_data.p, n_rows,
                                                                                   n_cols, X.p);
            }
            PCA_CU(cudaGetLastError());
            return pca_device(X.p, n_rows, n_cols, n_components, s, ev, out, sv, comp, mean);
        };
        rc = body();
    }
    if (rc == SCHIC_OK) {
        float total = 0.f;
        for (int i = 0; i < 6; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev.ev[i], ev.ev[i + 1]) != cudaSuccess) { cudaGetLastError(); ms = 0.f; }
            g_stage_ms[i] = ms;
            total += ms;
        }
        g_stage_ms[6] = total;
    }
    return rc;
}

}  // namespace

extern "C" {

int schic_pca_fit_transform(const double* X, int64_t n_rows, int64_t n_cols, int32_t n_components, int32_t device,
                            double* out, double* singular_values, double* components, double* mean) {
    if (!X) return schic::set_error(SCHIC_ERR_INVALID_ARG, "null matrix");
    return pca_entry(X, n_rows, n_cols, nullptr, nullptr, nullptr, n_components, device, out, singular_values, components, mean);
}

int schic_pca_fit_transform_csr(int64_t n_rows, int64_t n_cols, const int64_t* indptr, const int32_t* indices,
                                const float* data, int32_t n_components, int32_t device, double* out,
                                double* singular_values, double* components, double* mean) {
    if (!indptr) return schic::set_error(SCHIC_ERR_INVALID_ARG, "null indptr");
    return pca_entry(nullptr, n_rows, n_cols, indptr, indices, data, n_components, device, out, singular_values, components, mean);
}

int schic_pca_stage_ms(float* out, int32_t n_out) {
    if (!out || n_out < 7) return schic::set_error(SCHIC_ERR_INVALID_ARG, "out must hold 7 floats");
    std::memcpy(out, g_stage_ms, 7 * sizeof(float));
    return SCHIC_OK;
}

}  // extern "C"
