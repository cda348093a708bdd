// Posterior kernels around the triangular products (sm_100a, FP64).
//
//   xcov_kernel       k(X*, T) per objective, never in candidate-major form: writes the training-index-major
//                     slab KsT[j][c] (and Gr[j][c] = c1 * (dphi/dr)/r for the gradient), and reduces the mean
//                     mu = k . alpha and the two-GEMM-identity sums of d mu/dx in registers
//                     (replaces gp.py:148-151 `gp.kernel_(X, X_train_)`, `K.dot(alpha_)` and gp.py:173-196).
//   gradsum_kernel    after v = K^-1 k*: S0' = sum_j Gr_j v_j and S'_k = sum_j Gr_j v_j ts_jk
//                     (replaces gp.py:201-205 without the (C,N,d) tensors).
//   epilogue_kernel   combines partial sums, sigma^2 = (c1+c2) - |w|^2 clamp sqrt (gp.py:159-167),
//                     d sigma = dvar / (2 sigma) (gp.py:206), un-normalises (base.py:106-109) and applies
//                     the acquisition transform (identity.py / ucb.py / ei.py / pi.py) — fused.
//
// Thread <-> candidate mapping everywhere: candidate coordinates and all per-candidate accumulators live in
// registers, training points are broadcast from shared memory, every global access is coalesced over c.
#pragma once
#include "common.cuh"

namespace aoed {

constexpr int XC_THREADS = 128;  // candidates per CTA
constexpr int XC_JT = 32;        // training points staged per shared-memory tile

struct ModelDev {            // per-objective constants, device array of n_obj of these
    double c1, c2;
    double y_mean, y_scale;
};

struct XcovArgs {
    const double* XT;        // [DP][ldc]  candidate chunk, dimension-major (padded dims are zero)
    const double* Ts;        // [m][Npad][DP] training points divided by the length scales
    const double* alpha;     // [m][Npad]
    const double* ell;       // [m][DP] (padded with 1)
    const ModelDev* model;   // [m]
    double* KsT;             // [m][Npad][ldc] or nullptr
    double* Gr;              // [m][Npad][ldc] or nullptr
    double* part;            // [m][JS][2+DP][ldc]
    int64_t strideSlab;      // Npad*ldc
    int N, Npad, ldc, JS, jchunk;
};

#ifndef AOED_XCOV_MINBLOCKS
#define AOED_XCOV_MINBLOCKS 2
#endif
#ifndef AOED_XCOV_UNROLL
#define AOED_XCOV_UNROLL 1
#endif
#ifndef AOED_XCOV_XS_SMEM
#define AOED_XCOV_XS_SMEM 0
#endif

// DE <= DP: the dimensions the loops run over.  DP fixes the row stride of Ts / the partial-sum layout (8, 16, 24, 32, ...);
// with n_var <= DP - 4 the instance DE = DP - 4 skips the zero padding (n_var = 20: 17 % of the distance and moment FMAs and
// 8 registers per thread at DP = 24).
template <int NU, int DP, bool GRAD, int DE = DP>
__global__ void __launch_bounds__(XC_THREADS, AOED_XCOV_MINBLOCKS) xcov_kernel(XcovArgs p) {
    __shared__ __align__(16) double ts[XC_JT][DP];
    __shared__ double al[XC_JT];
    const int tid = threadIdx.x;
    const int c = blockIdx.x * XC_THREADS + tid;
    const int js = blockIdx.y, o = blockIdx.z;
    const double c1 = p.model[o].c1, c2 = p.model[o].c2;
    const double* __restrict__ Ts = p.Ts + int64_t(o) * p.Npad * DP;
    const double* __restrict__ alpha = p.alpha + int64_t(o) * p.Npad;

#if AOED_XCOV_XS_SMEM
    // candidate coordinates in shared memory ([k][thread]: conflict free) instead of 2*DP registers: more resident warps
    __shared__ double xsm[DE][XC_THREADS];
#pragma unroll
    for (int k = 0; k < DE; ++k) xsm[k][tid] = p.XT[int64_t(k) * p.ldc + c] / p.ell[o * DP + k];
#define xs(k) xsm[k][tid]
#else
    double xsr[DE];
#pragma unroll
    for (int k = 0; k < DE; ++k) xsr[k] = p.XT[int64_t(k) * p.ldc + c] / p.ell[o * DP + k];
#define xs(k) xsr[k]
#endif

    double mu = 0.0, s0 = 0.0;
    double sk[GRAD ? DE : 1];
    if (GRAD) {
#pragma unroll
        for (int k = 0; k < DE; ++k) sk[k] = 0.0;
    }
    double* __restrict__ KsT = p.KsT ? p.KsT + o * p.strideSlab : nullptr;
    double* __restrict__ Gr = p.Gr ? p.Gr + o * p.strideSlab : nullptr;

    const int j_begin = js * p.jchunk;
    const int j_end = min(p.N, j_begin + p.jchunk);
    for (int j0 = j_begin; j0 < j_end; j0 += XC_JT) {
        const int nj = min(XC_JT, j_end - j0);
        __syncthreads();
        for (int i = tid; i < nj * DP; i += XC_THREADS) (&ts[0][0])[i] = Ts[int64_t(j0) * DP + i];
        if (tid < nj) al[tid] = alpha[j0 + tid];
        __syncthreads();
#if AOED_XCOV_UNROLL == 1
        for (int jj = 0; jj < nj; ++jj) {
            double r2 = 0.0, r2b = 0.0;
#pragma unroll
            for (int k = 0; k < DE; k += 2) {
                const double2 t = *reinterpret_cast<const double2*>(&ts[jj][k]);
                const double d0 = xs(k) - t.x, d1 = xs(k + 1) - t.y;
                r2 = fma(d0, d0, r2);
                r2b = fma(d1, d1, r2b);
            }
            r2 += r2b;
            const double r = sqrt(r2);
            double phi, gor;
            kernel_profile<NU>(r, r2, phi, gor);
            const double kv = fma(c1, phi, c2);
            const double a = al[jj];
            mu = fma(kv, a, mu);
            if (KsT) KsT[int64_t(j0 + jj) * p.ldc + c] = kv;
            if (GRAD) {
                const double w = c1 * gor;
                if (Gr) Gr[int64_t(j0 + jj) * p.ldc + c] = w;
                const double wa = w * a;
                s0 += wa;
#pragma unroll
                for (int k = 0; k < DE; k += 2) {
                    const double2 t = *reinterpret_cast<const double2*>(&ts[jj][k]);
                    sk[k] = fma(wa, t.x, sk[k]);
                    sk[k + 1] = fma(wa, t.y, sk[k + 1]);
                }
            }
        }
    }
#else
        // two training points per iteration with split distance accumulators: the serial FMA -> sqrt -> exp chains of
        // independent points overlap (the kernel is FP64-latency bound at 4-8 warps per SM)
        for (int jj = 0; jj < nj; jj += 2) {
            const bool two = jj + 1 < nj;
            const int j1 = two ? jj + 1 : jj;
            double ra0 = 0.0, ra1 = 0.0, rb0 = 0.0, rb1 = 0.0;
#pragma unroll
            for (int k = 0; k < DE; k += 2) {
                const double2 ta = *reinterpret_cast<const double2*>(&ts[jj][k]);
                const double2 tb = *reinterpret_cast<const double2*>(&ts[j1][k]);
                const double a0 = xs(k) - ta.x, a1 = xs(k + 1) - ta.y;
                const double b0 = xs(k) - tb.x, b1 = xs(k + 1) - tb.y;
                ra0 = fma(a0, a0, ra0);
                ra1 = fma(a1, a1, ra1);
                rb0 = fma(b0, b0, rb0);
                rb1 = fma(b1, b1, rb1);
            }
            const double r2a = ra0 + ra1, r2b = rb0 + rb1;
            const double ra = sqrt(r2a), rb = sqrt(r2b);
            double phia, gora, phib, gorb;
            kernel_profile<NU>(ra, r2a, phia, gora);
            kernel_profile<NU>(rb, r2b, phib, gorb);
            const double kva = fma(c1, phia, c2), kvb = fma(c1, phib, c2);
            const double aa = al[jj], ab = two ? al[j1] : 0.0;
            mu = fma(kva, aa, mu);
            mu = fma(kvb, ab, mu);
            if (KsT) {
                KsT[int64_t(j0 + jj) * p.ldc + c] = kva;
                if (two) KsT[int64_t(j0 + j1) * p.ldc + c] = kvb;
            }
            if (GRAD) {
                const double wa = c1 * gora, wb = c1 * gorb;
                if (Gr) {
                    Gr[int64_t(j0 + jj) * p.ldc + c] = wa;
                    if (two) Gr[int64_t(j0 + j1) * p.ldc + c] = wb;
                }
                const double waa = wa * aa, wab = wb * ab;
                s0 += waa;
                s0 += wab;
#pragma unroll
                for (int k = 0; k < DE; k += 2) {
                    const double2 ta = *reinterpret_cast<const double2*>(&ts[jj][k]);
                    const double2 tb = *reinterpret_cast<const double2*>(&ts[j1][k]);
                    sk[k] = fma(waa, ta.x, sk[k]);
                    sk[k + 1] = fma(waa, ta.y, sk[k + 1]);
                    sk[k] = fma(wab, tb.x, sk[k]);
                    sk[k + 1] = fma(wab, tb.y, sk[k + 1]);
                }
            }
        }
    }
#endif
#undef xs
    double* __restrict__ part = p.part + (int64_t(o) * p.JS + js) * (2 + DP) * p.ldc + c;
    part[0] = mu;
    if (GRAD) {
        part[p.ldc] = s0;
#pragma unroll
        for (int k = 0; k < DE; ++k) part[int64_t(2 + k) * p.ldc] = sk[k];
    }
}

struct GradsumArgs {
    const double* V;    // [m][Npad][ldc]
    const double* Gr;   // [m][Npad][ldc]
    const double* Ts;   // [m][Npad][DP]
    double* part2;      // [m][JS][1+DP][ldc]
    int64_t strideSlab;
    int N, Npad, ldc, JS, jchunk;
};

constexpr int GS_B = 8;  // training points per register batch

template <int DP, int DE = DP>
__global__ void __launch_bounds__(XC_THREADS) gradsum_kernel(GradsumArgs p) {
    __shared__ __align__(16) double ts[XC_JT][DP];
    const int tid = threadIdx.x;
    const int c = blockIdx.x * XC_THREADS + tid;
    const int js = blockIdx.y, o = blockIdx.z;
    const double* __restrict__ Ts = p.Ts + int64_t(o) * p.Npad * DP;
    const double* __restrict__ V = p.V + o * p.strideSlab + c;
    const double* __restrict__ Gr = p.Gr + o * p.strideSlab + c;
    double s0 = 0.0, sk[DE];
#pragma unroll
    for (int k = 0; k < DE; ++k) sk[k] = 0.0;
    const int j_begin = js * p.jchunk;
    const int j_end = min(p.N, j_begin + p.jchunk);
    // The kernel is HBM bound (two 8-byte streams per training point against 1 + DE FMAs): the loads of the NEXT eight
    // points are issued before the current eight are used, so a warp always has a batch in flight while it computes.
    auto fetch = [&](int jb, double (&buf)[GS_B]) {
#pragma unroll
        for (int u = 0; u < GS_B; ++u) {
            const int j = min(jb + u, j_end - 1);
            const double v = Gr[int64_t(j) * p.ldc] * V[int64_t(j) * p.ldc];
            buf[u] = (jb + u < j_end) ? v : 0.0;
        }
    };
    double cur[GS_B], nxt[GS_B];
    if (j_begin < j_end) fetch(j_begin, cur);
    for (int j0 = j_begin; j0 < j_end; j0 += XC_JT) {
        const int nj = min(XC_JT, j_end - j0);
        __syncthreads();
        for (int i = tid; i < nj * DP; i += XC_THREADS) (&ts[0][0])[i] = Ts[int64_t(j0) * DP + i];
        __syncthreads();
#pragma unroll
        for (int b = 0; b < XC_JT / GS_B; ++b) {
            const int jb = b * GS_B;
            if (jb < nj) {
                if (j0 + jb + GS_B < j_end) fetch(j0 + jb + GS_B, nxt);
#pragma unroll
                for (int u = 0; u < GS_B; ++u) {
                    const int jj = min(jb + u, nj - 1);  // past the end: cur[u] == 0
                    s0 += cur[u];
#pragma unroll
                    for (int k = 0; k < DE; k += 2) {
                        double2 t = *reinterpret_cast<const double2*>(&ts[jj][k]);
                        sk[k] = fma(cur[u], t.x, sk[k]);
                        sk[k + 1] = fma(cur[u], t.y, sk[k + 1]);
                    }
                }
#pragma unroll
                for (int u = 0; u < GS_B; ++u) cur[u] = nxt[u];
            }
        }
    }
    double* __restrict__ part = p.part2 + (int64_t(o) * p.JS + js) * (1 + DP) * p.ldc + c;
    part[0] = s0;
#pragma unroll
    for (int k = 0; k < DE; ++k) part[int64_t(1 + k) * p.ldc] = sk[k];
}

// Transposes the candidate chunk [rows][d] (row-major, as numpy hands it over) to [DP][ldc].
static __global__ void transpose_x_kernel(const double* __restrict__ X, int64_t rows, int d, int DP, int ldc,
                                   double* __restrict__ XT) {
    const int64_t c = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (c >= ldc) return;
    for (int k = 0; k < DP; ++k) XT[int64_t(k) * ldc + c] = (c < rows && k < d) ? X[c * d + k] : 0.0;
}

struct AcqDev {
    int kind;         // AOED_ACQ_*
    int exact;        // AOED_EVAL_EXACT
    double lam;       // UCB: sqrt(log n / n)   (ucb.py:25)
    double y_min[16]; // EI / PI                (ei.py:21, pi.py:21)
};

struct EpilogueArgs {
    const double* XT;     // [DP][ldc]
    const double* ell;    // [m][DP]
    const ModelDev* model;
    const double* part;   // [m][JS][2+DP][ldc]
    const double* part2;  // [m][JS][1+DP][ldc]
    const double* sumsq;  // [m][mTiles][ldc]
    double *F, *S, *dF, *dS, *A, *dA;  // outputs for rows [0, rows): (rows, m) / (rows, m, d); any may be null
    int64_t rows;
    int m, d, DP, ldc, JS, JS2, mTiles;  // JS / JS2: training-index splits of xcov_kernel / gradsum_kernel
    int std, grad;
    AcqDev acq;
};

// value + first derivative of the acquisition in (mu, sigma)   (ucb.py:25-41, ei.py:27-46, pi.py:27-45)
__device__ __forceinline__ void acq_first_order(const AcqDev& acq, int o, double mu, double sd, double& val,
                                                double& a_mu, double& a_sd) {
    if (acq.kind == 1) {  // identity
        val = mu; a_mu = 1.0; a_sd = 0.0;
    } else if (acq.kind == 2) {  // UCB
        val = mu - acq.lam * sd; a_mu = 1.0; a_sd = -acq.lam;
    } else {
        const double diff = acq.y_min[o] - mu;
        const double z = safe_div(diff, sd);
        const double pdf = norm_pdf(z), cdf = norm_cdf(z);
        const double dz_m = -safe_div(1.0, sd);
        const double dz_s = -safe_div(diff, sd * sd);
        if (acq.kind == 3) {  // EI
            val = -diff * cdf - sd * pdf;
            if (acq.exact) {
                a_mu = cdf; a_sd = -pdf;
            } else {  // literal ei.py:40-41 (the sigma term has the reference's sign)
                const double dpdf = -z * pdf;
                a_mu = cdf - diff * pdf * dz_m - sd * dpdf * dz_m;
                a_sd = diff * pdf * dz_s + pdf + sd * dpdf * dz_s;
            }
        } else {  // PI
            val = -cdf;
            a_mu = -pdf * dz_m;
            a_sd = -pdf * dz_s;
        }
    }
}

constexpr int EPI_CX = 32;  // candidates per CTA
constexpr int EPI_KY = 8;   // threads per candidate: the partial-sum reads (~470 per candidate and objective) are split
                            // over them, otherwise the few thousand candidates of a chunk leave the GPU latency bound

static __global__ void __launch_bounds__(EPI_CX * EPI_KY) epilogue_kernel(EpilogueArgs p) {
    __shared__ double sc[4][EPI_CX];  // mu, sum of squares, s0, t0 per candidate
    const int cx = threadIdx.x, ky = threadIdx.y;
    const int64_t c = blockIdx.x * int64_t(EPI_CX) + cx;
    const int o = blockIdx.y;
    const bool live = c < p.rows;
    const int64_t cc = live ? c : 0;  // dead lanes read a valid column and store nothing
    const ModelDev md = p.model[o];
    const int ldc = p.ldc, DP = p.DP;
    const int64_t pstride = int64_t(2 + DP) * ldc;
    const double* part = p.part + int64_t(o) * p.JS * pstride + cc;
    const int64_t p2stride = int64_t(1 + DP) * ldc;
    const double* part2 = p.part2 + int64_t(o) * p.JS2 * p2stride + cc;
    // ---- phase 1: the four per-candidate scalars, one per ky (the sum of squares is shared by ky = 1, 4..7) ----
    if (ky == 0) {
        double mu = 0.0;
#pragma unroll 4
        for (int js = 0; js < p.JS; ++js) mu += part[js * pstride];
        sc[0][cx] = mu;
    } else if (ky == 2) {
        double s0 = 0.0;
        if (p.grad) {
#pragma unroll 4
            for (int js = 0; js < p.JS; ++js) s0 += part[js * pstride + ldc];
        }
        sc[2][cx] = s0;
    } else if (ky == 3) {
        double t0 = 0.0;
        if (p.grad && p.std) {
#pragma unroll 4
            for (int js = 0; js < p.JS2; ++js) t0 += part2[js * p2stride];
        }
        sc[3][cx] = t0;
    }
    __shared__ double ssq[5][EPI_CX];
    if (ky == 1 || ky >= 4) {
        const int slot = ky == 1 ? 0 : ky - 3;  // 5 readers, fixed tile ranges -> deterministic summation order
        double acc = 0.0;
        if (p.std) {
            const double* ss = p.sumsq + int64_t(o) * p.mTiles * ldc + cc;
            const int per = (p.mTiles + 4) / 5;
            const int t1 = min(p.mTiles, (slot + 1) * per);
#pragma unroll 4
            for (int t = slot * per; t < t1; ++t) acc += ss[int64_t(t) * ldc];
        }
        ssq[slot][cx] = acc;
    }
    __syncthreads();
    const double mu = sc[0][cx];
    double sd = 0.0;
    if (p.std) {
        const double acc = (((ssq[0][cx] + ssq[1][cx]) + ssq[2][cx]) + ssq[3][cx]) + ssq[4][cx];
        double var = (md.c1 + md.c2) - acc;  // kernel_.diag(X) - |L^-1 k*|^2   (gp.py:159-161)
        if (var < 0.0) var = 0.0;            // gp.py:163-165
        sd = sqrt(var);
    }
    const double F = fma(mu, md.y_scale, md.y_mean);  // base.py:106
    const double S = sd * md.y_scale;                 // base.py:108
    const int64_t io = cc * p.m + o;
    double aval = 0.0, a_mu = 1.0, a_sd = 0.0;
    if (p.acq.kind != 0) acq_first_order(p.acq, o, F, S, aval, a_mu, a_sd);
    if (ky == 0 && live) {
        if (p.F) p.F[io] = F;
        if (p.S && p.std) p.S[io] = S;
        if (p.A && p.acq.kind != 0) p.A[io] = aval;
    }
    if (!p.grad) return;
    // ---- phase 2: dimensions k = ky, ky + 8, ... ----
    const double s0 = sc[2][cx], t0 = sc[3][cx];
    for (int k = ky; k < p.d; k += EPI_KY) {
        const double ell = p.ell[o * DP + k];
        const double xs = p.XT[int64_t(k) * ldc + cc] / ell;
        double sk = 0.0;
#pragma unroll 4
        for (int js = 0; js < p.JS; ++js) sk += part[js * pstride + int64_t(2 + k) * ldc];
        const double dmu = (s0 * xs - sk) / ell;  // gp.py:196 via SURVEY.md A-4
        const double dFk = dmu * md.y_scale;      // base.py:107
        double dSk = 0.0;
        if (p.std) {
            double tk = 0.0;
#pragma unroll 4
            for (int js = 0; js < p.JS2; ++js) tk += part2[js * p2stride + int64_t(1 + k) * ldc];
            const double dvar = -2.0 * (t0 * xs - tk) / ell;  // gp.py:201-205
            dSk = 0.5 * safe_div(dvar, sd) * md.y_scale;      // gp.py:206, base.py:109
        }
        if (live) {
            const int64_t ig = io * p.d + k;
            if (p.dF) p.dF[ig] = dFk;
            if (p.dS && p.std) p.dS[ig] = dSk;
            if (p.dA && p.acq.kind != 0) p.dA[ig] = a_mu * dFk + a_sd * dSk;
        }
    }
}

}  // namespace aoed
