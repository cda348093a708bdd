// Fused small-batch posterior + acquisition VALUE kernel (sm_100a, FP64) for the MOBO regime: a few hundred to a few
// thousand candidates per call, n_train <= 256, no derivatives — what the NSGA-II generation loop asks for 200 times
// per proposal (solver/nsga2.py:17-30 -> surrogate_problem.py:32-52 -> acquisition -> gp.py:148-169).
//
// The chunked path (transpose, xcov, triangular product, epilogue: four kernels built for 10^4..10^6 candidates) takes
// ~40 us on 1000 candidates because each of its threads walks its share of the training set serially.  Here one CTA
// owns SE_TC = 16 candidates of one objective and spreads the work over (candidate, training point) pairs:
//   1. k*[j][c] for all pairs into shared memory, mean partials on the way        (gp.py:148-151)
//   2. w = L^-1 k*: thread i owns row i of L^-1 (read as column i of the stored L^-T: coalesced) for all 16 candidates,
//      k* broadcast from shared memory                                             (gp.py:159-161, triangular form)
//   3. |w|^2, sigma, un-normalisation, acquisition value                          (gp.py:163-167, base.py:106-109)
// Every candidate's result is independent of which tile it sits in, so the host call and the device-resident loop give
// identical numbers for the same design.
#pragma once
#include "posterior.cuh"

namespace aoed {

constexpr int SE_TC = 16;
constexpr int SE_THREADS = 256;
constexpr int SE_MAX_NPAD = 256;
constexpr int SE_MAX_ROWS = 4096;

struct SmallEvalArgs {
    const double* X;        // [rows][d] normalised candidates, row-major
    const double* Ts;       // [m][Npad][DP] training points divided by the length scales
    const double* alpha;    // [m][Npad]
    const double* ell;      // [m][DP]
    const ModelDev* model;  // [m]
    const double* LinvT;    // [m][Npad][Npad] row-major L^-T (zero below the diagonal and in the padding)
    int N, Npad, d, m, std;
    int64_t rows;
    AcqDev acq;
    double *F, *S, *A;      // [rows][m] (each may be null)
};

inline size_t small_eval_smem(int Npad, int DP) {
    return (size_t(Npad) * (SE_TC + 1) + size_t(Npad) * DP + size_t(Npad) + size_t(SE_TC) * (DP + 1) + 16 * SE_TC) * sizeof(double);
}
constexpr int SE_PF = 16;  // rows of L^-T in flight per thread: an L2 round trip is ~700 cycles, 16 rows of FMAs cover it

// sum over the 16 threads that share candidate c = tid % 16 (tid = c + 16 q): lanes l and l ^ 16 of a warp, then the 8
// warps in a fixed order -> deterministic
__device__ __forceinline__ double se_reduce(double v, double* red, int tid) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    __syncthreads();  // `red` may still be read from the previous use
    if ((tid & 31) < 16) red[(tid >> 5) * SE_TC + (tid & 15)] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < SE_THREADS / 32; ++w) s += red[w * SE_TC + (tid & 15)];
    return s;
}

template <int NU, int DP>
__global__ void __launch_bounds__(SE_THREADS) small_eval_kernel(SmallEvalArgs p) {
    extern __shared__ double se_smem[];
    constexpr int XS = DP + 1;                     // odd row stride: the 16 candidates of a warp hit 16 different banks
    constexpr int WS = SE_TC + 1;                  // row stride of the w buffer (same memory as ks, see step 2)
    double* ks = se_smem;                          // [Npad][SE_TC] k*, later [Npad][WS] w
    double* ts = ks + size_t(p.Npad) * WS;         // [Npad][DP] scaled training points of this objective
    double* al = ts + size_t(p.Npad) * DP;         // [Npad]
    double* xs = al + p.Npad;                      // [SE_TC][XS]
    double* red = xs + SE_TC * XS;                 // [8][SE_TC] (+ spare)
    const int tid = threadIdx.x, o = blockIdx.y;
    const int64_t c0 = int64_t(blockIdx.x) * SE_TC;
    const ModelDev md = p.model[o];
    {   // stage the model rows with all loads independent (read in place, every training point would cost an L2 round trip)
        const double* __restrict__ Ts = p.Ts + int64_t(o) * p.Npad * DP;
        const double* __restrict__ alpha = p.alpha + int64_t(o) * p.Npad;
        for (int t = tid; t < p.N * DP; t += SE_THREADS) ts[t] = Ts[t];
        for (int t = tid; t < p.N; t += SE_THREADS) al[t] = alpha[t];
    }
    for (int t = tid; t < SE_TC * DP; t += SE_THREADS) {
        const int c = t / DP, k = t - c * DP;
        const int64_t row = min(c0 + c, p.rows - 1);  // dead candidates of the last tile repeat a valid row
        xs[c * XS + k] = k < p.d ? p.X[row * p.d + k] / p.ell[o * DP + k] : 0.0;
    }
    __syncthreads();
    // ---- 1. cross covariances and the mean --------------------------------------------------------------
    const int c = tid & (SE_TC - 1);
    double mu = 0.0;
    for (int j = tid / SE_TC; j < p.N; j += SE_THREADS / SE_TC) {
        double r2 = 0.0, r2b = 0.0;
#pragma unroll
        for (int k = 0; k < DP; k += 2) {
            const double2 t = *reinterpret_cast<const double2*>(ts + j * DP + k);
            const double d0 = xs[c * XS + k] - t.x, d1 = xs[c * XS + k + 1] - t.y;
            r2 = fma(d0, d0, r2);
            r2b = fma(d1, d1, r2b);
        }
        r2 += r2b;
        double phi, gor;
        kernel_profile<NU>(sqrt(r2), r2, phi, gor);
        const double kv = fma(md.c1, phi, md.c2);
        ks[j * SE_TC + c] = kv;
        mu = fma(kv, al[j], mu);
    }
    mu = se_reduce(mu, red, tid);  // its barriers also publish ks
    // ---- 2. w = L^-1 k*, row i per thread ----------------------------------------------------------------
    double ss = 0.0;
    if (p.std) {
        double acc[SE_TC];
#pragma unroll
        for (int q = 0; q < SE_TC; ++q) acc[q] = 0.0;
        const int i = tid;
        if (i < p.N) {
            const double* __restrict__ col = p.LinvT + int64_t(o) * p.Npad * p.Npad + i;
            const int j_end = min(p.N, (i | 31) + 1);  // rows j > i of the column are exact zeros: no predicate needed
            double nxt[SE_PF];
#pragma unroll
            for (int u = 0; u < SE_PF; ++u) nxt[u] = col[int64_t(min(u, p.Npad - 1)) * p.Npad];
            for (int j0 = 0; j0 < j_end; j0 += SE_PF) {
                double cur[SE_PF];
#pragma unroll
                for (int u = 0; u < SE_PF; ++u) cur[u] = nxt[u];
#pragma unroll
                for (int u = 0; u < SE_PF; ++u)  // next batch (rows past the end are clamped: loaded, never used)
                    nxt[u] = col[int64_t(min(j0 + SE_PF + u, p.Npad - 1)) * p.Npad];
#pragma unroll
                for (int u = 0; u < SE_PF; ++u) {
                    const int j = j0 + u;
                    if (j < j_end) {
                        const double2* kj = reinterpret_cast<const double2*>(ks + j * SE_TC);
#pragma unroll
                        for (int q = 0; q < SE_TC / 2; ++q) {
                            const double2 kk = kj[q];
                            acc[2 * q] = fma(cur[u], kk.x, acc[2 * q]);
                            acc[2 * q + 1] = fma(cur[u], kk.y, acc[2 * q + 1]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone is done reading ks
        if (i < p.N) {  // odd row stride: a warp's 32 rows spread over the banks instead of all landing on one
#pragma unroll
            for (int q = 0; q < SE_TC; ++q) ks[i * WS + q] = acc[q];
        }
        __syncthreads();
        for (int r = tid / SE_TC; r < p.N; r += SE_THREADS / SE_TC) {
            const double w = ks[r * WS + c];
            ss = fma(w, w, ss);
        }
        ss = se_reduce(ss, red, tid);
    }
    // ---- 3. epilogue (same arithmetic as epilogue_kernel) -------------------------------------------------
    if (tid < SE_TC && c0 + tid < p.rows) {
        double sd = 0.0;
        if (p.std) {
            double var = (md.c1 + md.c2) - ss;  // gp.py:159-161
            if (var < 0.0) var = 0.0;           // gp.py:163-165
            sd = sqrt(var);
        }
        const double F = fma(mu, md.y_scale, md.y_mean);  // base.py:106
        const double S = sd * md.y_scale;                 // base.py:108
        const int64_t io = (c0 + tid) * p.m + o;
        if (p.F) p.F[io] = F;
        if (p.S && p.std) p.S[io] = S;
        if (p.A && p.acq.kind != 0) {
            double aval = 0.0, a_mu, a_sd;
            acq_first_order(p.acq, o, F, S, aval, a_mu, a_sd);
            p.A[io] = aval;
        }
    }
}

}  // namespace aoed
