// Kernels of the hyper-parameter fit (sm_100a, FP64): kernel matrix, LML gradient contraction, small helpers.
//
//   kmat_kernel       K = c1 * phi(r) + c2 (+ 1e-10 on the diagonal)      (gp.py:24-58 via sklearn kernel algebra,
//                     _gpr.py:582-589)
//   lml_grad_kernel   grad_theta = 1/2 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij/dtheta, fused: the (N,N,p) tensor of
//                     gp.py:63-79 / _gpr.py:627-650 is never materialised.
#pragma once
#include "common.cuh"

namespace aoed {

struct KmatArgs {
    const double* Ts;     // [B][Npad][DP] training points / length scale
    const double* c1c2;   // [B][2]
    double* K;            // [B][Npad][Npad], full symmetric
    int N, Npad, DP;
    int pad;              // 0: rows / columns >= N are left untouched; 1: identity padding up to Npad (blocked factorisation)
};

// Blocks of 64 x 64 pairs, 4 x 4 pairs per thread (rows ty + 16 r, columns tx + 16 c: coalesced stores, bank-conflict
// free shared-memory reads with rows padded to DP + 1).  Only blocks on or below the diagonal are computed: nothing
// downstream reads the upper triangle (the blocked Cholesky works on the lower one).
constexpr int KM_T = 64;
// coordinates staged per chunk of KC columns so that the two tiles stay below the 48 KB static shared-memory limit
template <int DP>
struct KmChunk {
    static constexpr int KC = DP <= 32 ? DP : (DP == 48 ? 24 : 32);
};

// stages columns [k0, k0 + KC) of the row tiles i0.. / j0.. (rows >= N read as 0)
template <int DP, int KC>
__device__ __forceinline__ void km_stage(const double* __restrict__ Ts, int N, int i0, int j0, int k0, double (*ti)[KC + 1],
                                         double (*tj)[KC + 1]) {
    for (int e = threadIdx.x; e < KM_T * KC; e += 256) {
        int r = e / KC, k = e % KC;
        ti[r][k] = (i0 + r < N) ? Ts[int64_t(i0 + r) * DP + k0 + k] : 0.0;
        tj[r][k] = (j0 + r < N) ? Ts[int64_t(j0 + r) * DP + k0 + k] : 0.0;
    }
}

template <int NU, int DP>
__global__ void __launch_bounds__(256) kmat_kernel(KmatArgs p) {
    constexpr int KC = KmChunk<DP>::KC;
    __shared__ double ti[KM_T][KC + 1];
    __shared__ double tj[KM_T][KC + 1];
    if (blockIdx.x > blockIdx.y) return;
    const int b = blockIdx.z;
    const int i0 = blockIdx.y * KM_T, j0 = blockIdx.x * KM_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const double* __restrict__ Ts = p.Ts + int64_t(b) * p.Npad * DP;
    double r2[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) r2[r][c] = 0.0;
    for (int k0 = 0; k0 < DP; k0 += KC) {
        if (k0 > 0) __syncthreads();
        km_stage<DP, KC>(Ts, p.N, i0, j0, k0, ti, tj);
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < KC; ++k) {
            double a[4], bb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                a[r] = ti[ty + 16 * r][k];
                bb[r] = tj[tx + 16 * r][k];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double d = a[r] - bb[c];
                    r2[r][c] = fma(d, d, r2[r][c]);
                }
        }
    }
    const double c1 = p.c1c2[2 * b], c2 = p.c1c2[2 * b + 1];
    double* __restrict__ K = p.K + int64_t(b) * p.Npad * p.Npad;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 16 * r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx + 16 * c;
            if (i >= p.N || j >= p.N) {
                if (p.pad && i < p.Npad && j < p.Npad) K[int64_t(i) * p.Npad + j] = (i == j) ? 1.0 : 0.0;
                continue;
            }
            double phi, gor;
            kernel_profile<NU>(sqrt(r2[r][c]), r2[r][c], phi, gor);
            double v = fma(c1, phi, c2);
            if (i == j) v = (c1 + c2) + kJitter;  // np.fill_diagonal(K, 1) (gp.py:55) then K[diag] += alpha (_gpr.py:589)
            K[int64_t(i) * p.Npad + j] = v;
        }
    }
}

// dphi/dlog l_k = h(r) * D_k, D_k = ((x_ik - x_jk)/l_k)^2   (gp.py:63-79; RBF: sklearn K * D_k)
template <int NU>
__device__ __forceinline__ double h_profile(double r, double r2, double phi) {
    if (NU == 1) return r != 0.0 ? phi / r : 0.0;  // safe_divide + isfinite guard (gp.py:68-70)
    if (NU == 3) return 3.0 * exp(-kSqrt3 * r);
    if (NU == 5) {
        double s = kSqrt5 * r;
        return 5.0 / 3.0 * (s + 1.0) * exp(-s);
    }
    return phi;
}

struct LmlGradArgs {
    const double* Ts;     // [B][Npad][DP]
    const double* c1c2;   // [B][2]
    const double* alpha;  // [B][Npad]
    const double* Kinv;   // [B][Npad][Npad] (row-major lower triangle valid)
    double* partial;      // [B][nBlocks][DP+2]
    int N, Npad, DP, nBlocksX, nBlocksY;
};

// One CTA per 64 x 64 block of (i, j) on or below the diagonal (the summand is symmetric: blocks below count twice),
// 4 x 4 pairs per thread: distances in a first sweep over the coordinates, the length-scale components in a second one
// (direct differences recomputed, as gp.py:63-79 forms them), every component reduced over the block once per 16 pairs
// of a thread.  The per-block partials are summed by lml_grad_reduce_kernel.  Output order: [c1, l_1..l_DP, c2].
template <int NU, int DP>
__global__ void __launch_bounds__(256) lml_grad_kernel(LmlGradArgs p) {
    constexpr int KC = KmChunk<DP>::KC;
    __shared__ double ti[KM_T][KC + 1];
    __shared__ double tj[KM_T][KC + 1];
    __shared__ double red[8][DP + 2];
    const int b = blockIdx.z;
    const int64_t blk = int64_t(blockIdx.y) * p.nBlocksX + blockIdx.x;
    double* __restrict__ part = p.partial + (int64_t(b) * p.nBlocksX * p.nBlocksY + blk) * (DP + 2);
    if (blockIdx.x > blockIdx.y) {
        if (threadIdx.x < DP + 2) part[threadIdx.x] = 0.0;
        return;
    }
    const int i0 = blockIdx.y * KM_T, j0 = blockIdx.x * KM_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* __restrict__ Ts = p.Ts + int64_t(b) * p.Npad * DP;
    double wh[4][4];  // first the squared distance, then w * c1 * h(r)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) wh[r][c] = 0.0;
    for (int k0 = 0; k0 < DP; k0 += KC) {
        if (k0 > 0) __syncthreads();
        km_stage<DP, KC>(Ts, p.N, i0, j0, k0, ti, tj);
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < KC; ++k) {
            double a[4], bb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                a[r] = ti[ty + 16 * r][k];
                bb[r] = tj[tx + 16 * r][k];
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double d = a[r] - bb[c];
                    wh[r][c] = fma(d, d, wh[r][c]);
                }
        }
    }
    const double c1 = p.c1c2[2 * b], c2 = p.c1c2[2 * b + 1];
    const double* __restrict__ alpha = p.alpha + int64_t(b) * p.Npad;
    const double* __restrict__ Kinv = p.Kinv + int64_t(b) * p.Npad * p.Npad;
    const double sym = (blockIdx.x < blockIdx.y) ? 2.0 : 1.0;
    double g_c1 = 0.0, g_c2 = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + 16 * r;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx + 16 * c;
            double whv = 0.0;
            if (i < p.N && j < p.N) {
                // only the row-major LOWER triangle of K^-1 is valid, read it symmetrically
                const double w = sym * (alpha[i] * alpha[j] - Kinv[int64_t(max(i, j)) * p.Npad + min(i, j)]);
                const double r2 = wh[r][c], rr = sqrt(r2);
                double phi, gor;
                kernel_profile<NU>(rr, r2, phi, gor);
                if (i == j) phi = 1.0;
                g_c1 = fma(w * c1, phi, g_c1);
                g_c2 = fma(w, c2, g_c2);
                whv = w * c1 * h_profile<NU>(rr, r2, phi);
            }
            wh[r][c] = whv;
        }
    }
    auto warp_sum = [&](double v) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        return v;
    };
    g_c1 = warp_sum(g_c1);
    g_c2 = warp_sum(g_c2);
    if (lane == 0) {
        red[warp][0] = g_c1;
        red[warp][DP + 1] = g_c2;
    }
    for (int k0 = (DP / KC - 1) * KC; k0 >= 0; k0 -= KC) {  // the last chunk is still staged
        if (k0 != (DP / KC - 1) * KC) {
            __syncthreads();
            km_stage<DP, KC>(Ts, p.N, i0, j0, k0, ti, tj);
            __syncthreads();
        }
#pragma unroll 2
        for (int k = 0; k < KC; ++k) {
            double a[4], bb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                a[r] = ti[ty + 16 * r][k];
                bb[r] = tj[tx + 16 * r][k];
            }
            double sk = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double d = a[r] - bb[c];
                    sk = fma(wh[r][c] * d, d, sk);
                }
            sk = warp_sum(sk);
            if (lane == 0) red[warp][1 + k0 + k] = sk;
        }
    }
    __syncthreads();
    if (threadIdx.x < DP + 2) {
        double s = 0.0;
        for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
        part[threadIdx.x] = s;
    }
}

// res[1 + k] = 0.5 * sum_blocks partial (k in the theta order [c1, l_1..l_d, c2]); zeros when the factorisation failed.
// One CTA per component k (deterministic tree reduction).
// Batched over blockIdx.y: problem b reads partial + b * nBlocks * (DP + 2), info[b], writes res + b * (d + 3).
static __global__ void __launch_bounds__(256) lml_grad_reduce_kernel(const double* __restrict__ partial, int nBlocks, int DP, int d,
                                                              const int* info, double* res) {
    __shared__ double red[8];
    const int k = blockIdx.x;
    partial += int64_t(blockIdx.y) * nBlocks * (DP + 2);
    info += blockIdx.y;
    res += int64_t(blockIdx.y) * (d + 3);
    double s = 0.0;
    for (int i = threadIdx.x; i < nBlocks; i += 256) s += partial[int64_t(i) * (DP + 2) + k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x != 0) return;
    s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    s = (*info != 0) ? 0.0 : 0.5 * s;
    if (k == 0) res[1] = s;
    else if (k == DP + 1) res[1 + d + 1] = s;
    else if (k - 1 < d) res[1 + k] = s;
}

// y_out[i] = sum_j A[i][j] x[j] over the stored triangle of a row-major matrix: LOWER (j <= i) or UPPER (j >= i).
// One warp per row; used for z = L^-1 y and alpha = L^-T z of the blocked LML path.
// Batched over blockIdx.y: matrix A + b * strideA, vector x + (xsel ? xsel[b] : b) * strideX, result y_out + b * strideY.
template <bool UPPER>
__global__ void __launch_bounds__(256) tri_matvec_kernel(const double* __restrict__ A, int64_t strideA, const double* __restrict__ x,
                                                         int64_t strideX, const int* __restrict__ xsel, int N, int ld,
                                                         double* __restrict__ y_out, int64_t strideY) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const int b = blockIdx.y;
    A += b * strideA;
    x += int64_t(xsel ? xsel[b] : b) * strideX;
    y_out += b * strideY;
    const double* a = A + int64_t(row) * ld;
    const int j0 = UPPER ? row : 0, j1 = UPPER ? N : row + 1;
    double s = 0.0;
    for (int j = j0 + lane; j < j1; j += 32) s = fma(a[j], x[j], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y_out[row] = s;
}

// Ts = X / l (padded to DP), c1c2 = exp(theta[0]), exp(theta[d+1]) for one hyper-parameter vector (device-side, no host sync)
// Batched over blockIdx.y (theta + b * (d + 2), Ts + b * Npad * DP, c1c2 + 2 b).
static __global__ void theta_prep_kernel(const double* __restrict__ X, const double* __restrict__ theta, int N, int Npad, int d,
                                  int DP, double* __restrict__ Ts, double* __restrict__ c1c2) {
    theta += int64_t(blockIdx.y) * (d + 2);
    Ts += int64_t(blockIdx.y) * Npad * DP;
    c1c2 += 2 * blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) {
        c1c2[0] = exp(theta[0]);
        c1c2[1] = exp(theta[d + 1]);
    }
    if (idx >= Npad * DP) return;
    const int i = idx / DP, k = idx % DP;
    Ts[idx] = (i < N && k < d) ? X[int64_t(i) * d + k] / exp(theta[1 + k]) : 0.0;
}

// res[0] = -1/2 y.alpha - sum log L_ii - N/2 log 2 pi   (_gpr.py:601-616); -inf when potrf reported a non-PD matrix
// quad_a . quad_b is y . alpha (sklearn's form) or z . z with z = L^-1 y (same number, better conditioned)
// Batched over blockIdx.x: y, alpha + b * Npad, Lfac + b * Npad^2, info[b], res + b * resStride.
static __global__ void __launch_bounds__(256) lml_value_kernel(const double* __restrict__ y, const double* __restrict__ alpha,
                                                        const double* __restrict__ Lfac, int N, int Npad, const int* info,
                                                        double* res, int resStride) {
    __shared__ double red[2][8];
    y += int64_t(blockIdx.x) * Npad;
    alpha += int64_t(blockIdx.x) * Npad;
    Lfac += int64_t(blockIdx.x) * Npad * Npad;
    info += blockIdx.x;
    res += int64_t(blockIdx.x) * resStride;
    double q = 0.0, ld = 0.0;
    for (int i = threadIdx.x; i < N; i += 256) {
        q = fma(y[i], alpha[i], q);
        ld += log(Lfac[int64_t(i) * Npad + i]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        q += __shfl_xor_sync(0xffffffffu, q, off);
        ld += __shfl_xor_sync(0xffffffffu, ld, off);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = q;
        red[1][threadIdx.x >> 5] = ld;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double qq = 0.0, ll = 0.0;
        for (int w = 0; w < 8; ++w) {
            qq += red[0][w];
            ll += red[1][w];
        }
        res[0] = (*info != 0) ? -INFINITY : -0.5 * qq - ll - 0.5 * N * 1.8378770664093453;
    }
}

// out[j][i] = in[i][j] for a square Npad matrix (batched via blockIdx.z)
static __global__ void transpose_square_kernel(const double* in, double* out, int n) {
    __shared__ double tile[32][33];
    const double* src = in + int64_t(blockIdx.z) * n * n;
    double* dst = out + int64_t(blockIdx.z) * n * n;
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
    for (int r = 0; r < 32; r += 8)
        if (x < n && y + r < n) tile[threadIdx.y + r][threadIdx.x] = src[int64_t(y + r) * n + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x;
    y = blockIdx.x * 32 + threadIdx.y;
    for (int r = 0; r < 32; r += 8)
        if (x < n && y + r < n) dst[int64_t(y + r) * n + x] = tile[threadIdx.x][threadIdx.y + r];
}

// keeps the lower triangle (row-major, rows/cols < N) and zeroes everything else
static __global__ void keep_lower_kernel(double* A, int N, int Npad) {
    double* M = A + int64_t(blockIdx.z) * Npad * Npad;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= Npad) return;
    if (i >= N || j >= N || j > i) M[int64_t(i) * Npad + j] = 0.0;
}

static __global__ void set_identity_kernel(double* A, int N, int Npad) {
    double* M = A + int64_t(blockIdx.z) * Npad * Npad;
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= Npad) return;
    M[int64_t(i) * Npad + j] = (i == j && i < N) ? 1.0 : 0.0;
}

}  // namespace aoed
