// Second-derivative kernels of the GP posterior (sm_100a, FP64).
//
// Replaces gp.py:209-242 (`hK`, `hy_mean`, `hy_var`, `hy_std`) and the second-order parts of the acquisitions
// (ucb.py:44-51, ei.py:48-68, pi.py:47-62) WITHOUT the (C, N, d, d) tensors the reference materialises.
//
// With s_i = (x_i - t_i)/l_i^2 every per-training-point Hessian of the kernel is
//     hK_ij(t) = A_t * delta_ij / l_i^2 + B_t * s_i s_j + C_t * s_j^2
// (C_t only in `literal` mode, where the reference's `dd ** 2` broadcast indexes the column j; in
// `exact` mode the q-term joins B_t), so  sum_t a_t hK_ij(t)  is one scalar, one d x d moment matrix and one d-vector
// per weight vector a (alpha for the mean, v = K^-1 k* for the variance).
//
//   dkcol_kernel   column slab DK[t][c*d+i] = d k(x_c, t)/d x_i   (gp.py:181-191) — the right-hand sides of the
//                  triangular product that gives (dK_i K^-1 dK_i^T) = |L^-1 dk_i|^2  (gp.py:236-240)
//   gram_kernel    exact mode only: (L^-1 dk_i) . (L^-1 dk_j)
//   hess_kernel    one CTA per (candidate, objective): moments, hF, hS and the acquisition Hessian, fused.
#pragma once
#include "common.cuh"
#include "posterior.cuh"

namespace aoed {

constexpr int HS_THREADS = 256;
constexpr int HS_TT = 128;  // training points per shared-memory tile

struct DkcolArgs {
    const double* XT;    // [DP][ldx] candidates, dimension-major
    const double* Ts;    // [m][Npad][DP]
    const double* ell;   // [m][DP]
    const double* Gr;    // [m][Npad][ldc]   c1 * (dphi/dr)/r per (training point, candidate)
    double* DK;          // [m][Npad][ldc]   columns c*d+i
    int64_t strideSlab;
    int rows, d, DP, N, Npad, ldc, ldx, JS, jchunk;
};

static __global__ void __launch_bounds__(128) dkcol_kernel(DkcolArgs p) {
    const int col = blockIdx.x * 128 + threadIdx.x;
    const int js = blockIdx.y, o = blockIdx.z;
    const int ncol = p.rows * p.d;
    const bool live = col < ncol;
    const int c = live ? col / p.d : 0, i = live ? col % p.d : 0;
    const double ell = p.ell[o * p.DP + i];
    const double xs = p.XT[int64_t(i) * p.ldx + c] / ell;
    const double* __restrict__ Ts = p.Ts + int64_t(o) * p.Npad * p.DP + i;
    const double* __restrict__ Gr = p.Gr + o * p.strideSlab + c;
    double* __restrict__ DK = p.DK + o * p.strideSlab + col;
    const int j_end = min(p.N, (js + 1) * p.jchunk);
    for (int j = js * p.jchunk; j < j_end; ++j) {
        // dk/dx_i = c1 (g/r) (x_i - t_i) / l_i^2      (gp.py:176-191 with dd = (x - t)/(r l^2))
        const double v = live ? Gr[int64_t(j) * p.ldc] * (xs - Ts[int64_t(j) * p.DP]) / ell : 0.0;
        DK[int64_t(j) * p.ldc] = v;
    }
}

struct GramArgs {
    const double* W;   // [m][Npad][ldc] = L^-1 DK, columns c*d+i
    double* G;         // [m][rows][d][d]
    int64_t strideSlab;
    int rows, d, N, ldc;
};

// one CTA per (candidate, objective); thread per (i, j) pair
constexpr int GR_MAX_D = 128;
constexpr int GR_TT = 32;  // rows of W per shared-memory tile (32 x 129 doubles stay inside 48 KB of static memory)
constexpr int GR_PPT = GR_MAX_D * GR_MAX_D / HS_THREADS;  // (i, j) pairs per thread at the largest d
static __global__ void __launch_bounds__(HS_THREADS) gram_kernel(GramArgs p) {
    __shared__ double w[GR_TT][GR_MAX_D + 1];
    const int c = blockIdx.x, o = blockIdx.y, d = p.d;
    const double* __restrict__ W = p.W + o * p.strideSlab + int64_t(c) * d;
    double acc[GR_PPT];
#pragma unroll
    for (int u = 0; u < GR_PPT; ++u) acc[u] = 0.0;
    for (int t0 = 0; t0 < p.N; t0 += GR_TT) {
        const int nt = min(GR_TT, p.N - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < nt * d; e += HS_THREADS) {
            const int t = e / d, i = e % d;
            w[t][i] = W[int64_t(t0 + t) * p.ldc + i];
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < GR_PPT; ++u) {
            const int pr = threadIdx.x + u * HS_THREADS;
            if (pr < d * d) {
                const int i = pr / d, j = pr % d;
                double a = acc[u];
                for (int t = 0; t < nt; ++t) a = fma(w[t][i], w[t][j], a);
                acc[u] = a;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < GR_PPT; ++u) {
        const int pr = threadIdx.x + u * HS_THREADS;
        if (pr < d * d) p.G[(int64_t(o) * p.rows + c) * d * d + pr] = acc[u];
    }
}

// radial coefficients of hK = P(r) * hd + Q(r) * q   (gp.py:216-226), already multiplied by c1
template <int NU>
__device__ __forceinline__ void hess_profile(double c1, double r, double& P, double& Q) {
    if (NU == 1) {
        const double e = c1 * exp(-r);
        P = -e;
        Q = e;
    } else if (NU == 3) {
        const double e = -3.0 * c1 * exp(-kSqrt3 * r);
        P = e * r;
        Q = e * (1.0 - kSqrt3 * r);
    } else if (NU == 5) {
        const double e = -5.0 / 3.0 * c1 * exp(-kSqrt5 * r);
        const double s = 1.0 + kSqrt5 * r;
        P = e * s * r;
        Q = e * (s - 5.0 * r * r);
    } else {
        const double e = -c1 * exp(-0.5 * r * r);
        P = e * r;
        Q = e * (1.0 - r * r);
    }
}

struct HessArgs {
    const double* XT;      // [DP][ldx]
    const double* Ts;      // [m][Npad][DP]
    const double* ell;     // [m][DP]
    const double* alpha;   // [m][Npad]
    const ModelDev* model; // [m]
    const double* V;       // [m][Npad][ldc] = K^-1 k*, column c   (std only)
    const double* sumsq;   // [m][mTiles][ldc]  |L^-1 k*|^2 partials, column c          (std only)
    const double* mid_ss;  // [m][mTiles][ldc]  |L^-1 dk_i|^2 partials, column c*d+i     (std, literal)
    const double* gram;    // [m][rows][d][d]   (L^-1 dk_i).(L^-1 dk_j)                  (std, exact)
    const double *F, *S, *dF, *dS;  // raw-unit first-order outputs of the same rows: (rows, m) / (rows, m, d)
    double *hF, *hS, *hA;           // (rows, m, d, d); any may be null
    int64_t strideSlab;
    int rows, m, d, N, Npad, ldc, ldx, mTiles;
    int std;
    AcqDev acq;
};

// second-order coefficients of the acquisition in (mu, sigma): A_ij = a hmu_ij + b hsd_ij + h_mm dmu_i dmu_j
//   + h_ms (dmu_i dsd_j + dsd_i dmu_j) + h_ss dsd_i dsd_j      (literal mode has h_ms = 0: ei.py:67-68, pi.py:61-62)
__device__ __forceinline__ void acq_second_order(const AcqDev& acq, int o, double mu, double sd, double& a, double& b,
                                                 double& h_mm, double& h_ms, double& h_ss) {
    h_mm = h_ms = h_ss = 0.0;
    if (acq.kind == 1) {
        a = 1.0; b = 0.0;
        return;
    }
    if (acq.kind == 2) {
        a = 1.0; b = -acq.lam;
        return;
    }
    const double diff = acq.y_min[o] - mu;
    const double z = safe_div(diff, sd);
    const double pdf = norm_pdf(z), cdf = norm_cdf(z);
    const double dz_m = -safe_div(1.0, sd);
    const double dz_s = -safe_div(diff, sd * sd);
    const double dpdf = -z * pdf;
    if (acq.kind == 3) {  // EI
        if (acq.exact) {
            const double inv = safe_div(1.0, sd);
            a = cdf; b = -pdf;
            h_mm = -pdf * inv; h_ms = -pdf * z * inv; h_ss = -z * z * pdf * inv;
        } else {  // ei.py:36-41, 49-60
            a = cdf - diff * pdf * dz_m - sd * dpdf * dz_m;
            b = diff * pdf * dz_s + pdf + sd * dpdf * dz_s;
            const double dpdf_m = dpdf * dz_m, dpdf_s = dpdf * dz_s;
            const double ddpdf_m = -z * dpdf_m - dz_m * pdf;
            const double ddpdf_s = -z * dpdf_s - dz_s * pdf;
            const double ddz_ss = safe_div(diff, sd * sd * sd);
            h_mm = -pdf * dz_m - dz_m * pdf + diff * dpdf * dz_m * dz_m + sd * dz_m * ddpdf_m;
            h_ss = diff * (dz_s * dpdf_s + pdf * ddz_ss) + dpdf_s + dpdf * dz_s + sd * dz_s * ddpdf_s + sd * dpdf * ddz_ss;
        }
    } else {  // PI
        a = -pdf * dz_m;
        b = -pdf * dz_s;
        if (acq.exact) {
            const double inv2 = safe_div(1.0, sd * sd);
            h_mm = z * pdf * inv2; h_ms = (z * z - 1.0) * pdf * inv2; h_ss = (z * z * z - 2.0 * z) * pdf * inv2;
        } else {  // pi.py:48-55
            const double hz_s = safe_div(diff, sd * sd * sd);
            h_mm = -(dpdf * dz_m) * dz_m;
            h_ss = -(dpdf * dz_s) * dz_s - pdf * hz_s;
        }
    }
}

template <int NU, int DP>
__global__ void __launch_bounds__(HS_THREADS) hess_kernel(HessArgs p) {
    constexpr int PPT = (DP * DP + HS_THREADS - 1) / HS_THREADS;  // (i, j) pairs per thread
    constexpr int TT = DP > 64 ? HS_TT / 4 : (DP > 32 ? HS_TT / 2 : HS_TT);  // the tile of s vectors has to stay inside 48 KB of static memory
    __shared__ double sS[TT][DP + 1];   // s_t[i] = (x_i - t_i) / l_i^2
    __shared__ double wB[2][TT], wC[2][TT];
    __shared__ double xs[DP], il[DP];
    __shared__ double redA[2][HS_THREADS / 32];
    const int tid = threadIdx.x;
    const int c = blockIdx.x, o = blockIdx.y, d = p.d;
    const bool exact = p.acq.exact != 0;
    const ModelDev md = p.model[o];
    const double* __restrict__ Ts = p.Ts + int64_t(o) * p.Npad * DP;
    const double* __restrict__ alpha = p.alpha + int64_t(o) * p.Npad;
    const double* __restrict__ V = p.std ? p.V + o * p.strideSlab + c : nullptr;
    if (tid < DP) {
        const double l = p.ell[o * DP + tid];
        il[tid] = 1.0 / l;
        xs[tid] = p.XT[int64_t(tid) * p.ldx + c] / l;
    }
    double accB[2][PPT], accC[2][PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) accB[0][u] = accB[1][u] = accC[0][u] = accC[1][u] = 0.0;
    double sumA[2] = {0.0, 0.0};
    __syncthreads();

    for (int t0 = 0; t0 < p.N; t0 += TT) {
        const int nt = min(TT, p.N - t0);
        if (tid < TT) {
            double cA = 0.0, cB = 0.0, cC = 0.0, al = 0.0, vv = 0.0;
            if (tid < nt) {
                const double* tr = Ts + int64_t(t0 + tid) * DP;
                double r2 = 0.0;
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    const double dk = xs[k] - tr[k];
                    r2 = fma(dk, dk, r2);
                    sS[tid][k] = dk * il[k];
                }
                const double r = sqrt(r2);
                if (r != 0.0) {  // safe_divide: hd = dd = 0 at r = 0 (gp.py:176-179, 212-214)
                    double P, Q;
                    hess_profile<NU>(md.c1, r, P, Q);
                    const double ir = 1.0 / r, ir2 = ir * ir;
                    cA = P * ir;             // hd_ij = delta_ij/(r l_i^2) - s_i s_j / r^3
                    cB = -P * ir2 * ir;
                    if (exact) cB += Q * ir2;  // q_ij = u_i u_j = s_i s_j / r^2
                    else cC = Q * ir2;         // q_j  = u_j^2   = s_j^2 / r^2   
                }
                al = alpha[t0 + tid];
                if (p.std) vv = V[int64_t(t0 + tid) * p.ldc];
            } else {
#pragma unroll
                for (int k = 0; k < DP; ++k) sS[tid][k] = 0.0;
            }
            wB[0][tid] = cB * al; wB[1][tid] = cB * vv;
            wC[0][tid] = cC * al; wC[1][tid] = cC * vv;
            sumA[0] = fma(cA, al, sumA[0]);
            sumA[1] = fma(cA, vv, sumA[1]);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int pr = tid + u * HS_THREADS;
            if (pr < d * d) {
                const int i = pr / d, j = pr % d;
                double b0 = accB[0][u], b1 = accB[1][u], c0 = accC[0][u], c1 = accC[1][u];
                for (int t = 0; t < TT; ++t) {
                    const double sj = sS[t][j];
                    const double sij = sS[t][i] * sj, sjj = sj * sj;
                    b0 = fma(wB[0][t], sij, b0);
                    b1 = fma(wB[1][t], sij, b1);
                    c0 = fma(wC[0][t], sjj, c0);
                    c1 = fma(wC[1][t], sjj, c1);
                }
                accB[0][u] = b0; accB[1][u] = b1; accC[0][u] = c0; accC[1][u] = c1;
            }
        }
        __syncthreads();
    }
    // block sums of A_t a_t
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        double v = sumA[w];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((tid & 31) == 0) redA[w][tid >> 5] = v;
    }
    __syncthreads();
    double totA[2] = {0.0, 0.0};
#pragma unroll
    for (int w = 0; w < 2; ++w)
        for (int k = 0; k < HS_THREADS / 32; ++k) totA[w] += redA[w][k];

    // per-candidate scalars
    const int64_t io = int64_t(c) * p.m + o;
    double var = 0.0, sd = 0.0;
    if (p.std) {
        double acc = 0.0;
        const double* ss = p.sumsq + int64_t(o) * p.mTiles * p.ldc + c;
        for (int t = 0; t < p.mTiles; ++t) acc += ss[int64_t(t) * p.ldc];
        var = (md.c1 + md.c2) - acc;  // gp.py:159-165
        if (var < 0.0) var = 0.0;
        sd = sqrt(var);
    }
    const double out_scale = exact ? md.y_scale : 1.0;  // literal: hF, hS stay in normalised-Y units
    double a = 1.0, b = 0.0, h_mm = 0.0, h_ms = 0.0, h_ss = 0.0;
    if (p.acq.kind != 0 && p.hA)
        acq_second_order(p.acq, o, p.F[io], p.std ? p.S[io] : 0.0, a, b, h_mm, h_ms, h_ss);
    const double inv_ys = 1.0 / md.y_scale;
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        const int pr = tid + u * HS_THREADS;
        if (pr >= d * d) continue;
        const int i = pr / d, j = pr % d;
        const double diag = (i == j) ? il[i] * il[i] : 0.0;
        const double hmu = fma(diag, totA[0], accB[0][u] + accC[0][u]);  // gp.py:230
        const int64_t ih = io * d * d + pr;
        if (p.hF) p.hF[ih] = hmu * out_scale;
        double hsd = 0.0;
        if (p.std) {
            const double hkv = fma(diag, totA[1], accB[1][u] + accC[1][u]);  // sum_t hK_ij,t v_t
            double mid;
            if (exact) {
                mid = p.gram[(int64_t(o) * p.rows + c) * d * d + pr];
            } else {  // 2 dK_i K^-1 dK_i^T: index i on both factors (gp.py:236-240)
                mid = 0.0;
                const double* ms = p.mid_ss + int64_t(o) * p.mTiles * p.ldc + int64_t(c) * d + i;
                for (int t = 0; t < p.mTiles; ++t) mid += ms[int64_t(t) * p.ldc];
            }
            const double hvar = -2.0 * (hkv + mid);  // gp.py:240
            // dy_var = 2 sd dy_std (gp.py:206), in normalised-Y units
            const double dsd_j = p.dS[io * d + j] * inv_ys;
            const double dsd_i = p.dS[io * d + i] * inv_ys;
            const double dvar = 2.0 * sd * (exact ? dsd_i : dsd_j);
            hsd = 0.5 * safe_div(hvar * sd - dvar * dsd_j, var);  // gp.py:241
            if (p.hS) p.hS[ih] = hsd * out_scale;
        }
        if (p.hA && p.acq.kind != 0) {
            const double dmi = p.dF[io * d + i], dmj = p.dF[io * d + j];
            double v = a * (hmu * out_scale) + b * (hsd * out_scale) + h_mm * dmi * dmj;
            if (p.std) {
                const double dsi = p.dS[io * d + i], dsj = p.dS[io * d + j];
                v += h_ss * dsi * dsj + h_ms * (dmi * dsj + dsi * dmj);
            }
            p.hA[ih] = v;
        }
    }
}

}  // namespace aoed
