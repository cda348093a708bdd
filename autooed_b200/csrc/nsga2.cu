// Device-resident NSGA-II generation loop on the GP acquisition (sm_100a) — SURVEY.md §8 f1.
//
// Replaces, for the common unconstrained case, the pymoo 0.4.1 `NSGA2` + `minimize(..., ('n_gen', G))` loop that
// `autooed/mobo/solver/nsga2.py:17-30` drives (one `SurrogateProblem.evaluate` per generation): variation, duplicate
// elimination, acquisition evaluation, non-dominated sorting, crowding distance and survival all stay on the GPU; the
// host only enqueues ~8 launches per generation and never synchronises inside the loop.
//
// Operators follow pymoo's defaults: binary tournament on (rank, crowding distance), simulated
// binary crossover (prob 0.9, eta 15, per-variable exchange 0.5), polynomial mutation (eta 20, prob 1/n_var), offspring
// equal to an existing individual (on a 1e-8 grid) are discarded, rank-and-crowding survival.  Random numbers come from
// Philox (counter based: individual x generation), so a run is reproducible for a given seed but NOT seed-identical with
// numpy / pymoo.
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <string>
#include <vector>

#include "../../include/aoed_b200.h"
#include "common.cuh"
#include "rank.cuh"
#include "rff.cuh"

#include <functional>

extern "C" int aoed_detail_fail(int code, const char* msg);

namespace {

using namespace aoed;

int nfail(int code, const std::string& msg) { return aoed_detail_fail(code, msg.c_str()); }
#define NCU(call)                                                                                                    \
    do {                                                                                                             \
        cudaError_t e_ = (call);                                                                                     \
        if (e_ != cudaSuccess)                                                                                       \
            return nfail(AOED_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + \
                                            std::to_string(__LINE__));                                               \
    } while (0)
#define NCHK(expr)                      \
    do {                                \
        int rc_ = (expr);               \
        if (rc_ != AOED_OK) return rc_; \
    } while (0)

constexpr int NS_MAX_D = 32;
constexpr int NS_SORT_THREADS = 1024;

struct VarArgs {
    const double* X;      // [n_cur][d] current population (continuous)
    const int* rank;      // [n_cur]
    const double* crowd;  // [n_cur]
    const double *xl, *xu;
    double* Xoff;         // [P][d] offspring (continuous)
    int n_cur, P, d;
    unsigned long long seed;
    int gen;
    double pc, eta_c, eta_m;
};

__device__ __forceinline__ int tournament(curandStatePhilox4_32_10_t& st, const VarArgs& p) {
    const int a = min(int(curand_uniform_double(&st) * p.n_cur), p.n_cur - 1);
    const int b = min(int(curand_uniform_double(&st) * p.n_cur), p.n_cur - 1);
    const double coin = curand_uniform_double(&st);
    const int ra = p.rank[a], rb = p.rank[b];
    if (ra != rb) return ra < rb ? a : b;
    const double ca = p.crowd[a], cb = p.crowd[b];
    if (ca != cb) return ca > cb ? a : b;
    return coin < 0.5 ? a : b;
}

__device__ __forceinline__ double sbx_betaq(double u, double beta_bound, double eta) {
    const double alpha = 2.0 - pow(beta_bound, -(eta + 1.0));
    return (u <= 1.0 / alpha) ? pow(u * alpha, 1.0 / (eta + 1.0)) : pow(1.0 / (2.0 - u * alpha), 1.0 / (eta + 1.0));
}

__device__ __forceinline__ double poly_mutate(double x, double lo, double hi, double u, double eta) {
    const double span = hi > lo ? hi - lo : 1.0;
    const double d1 = (x - lo) / span, d2 = (hi - x) / span, pw = 1.0 / (eta + 1.0);
    double dq;
    if (u <= 0.5)
        dq = pow(2.0 * u + (1.0 - 2.0 * u) * pow(1.0 - d1, eta + 1.0), pw) - 1.0;
    else
        dq = 1.0 - pow(2.0 * (1.0 - u) + 2.0 * (u - 0.5) * pow(1.0 - d2, eta + 1.0), pw);
    return fmin(fmax(x + dq * span, lo), hi);
}

// one thread per offspring PAIR: two tournaments per parent, SBX, polynomial mutation
__global__ void __launch_bounds__(128) variation_kernel(VarArgs p) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * pair >= p.P) return;
    curandStatePhilox4_32_10_t st;
    curand_init(p.seed, (unsigned long long)pair, (unsigned long long)p.gen * 4096ull, &st);
    const int ia = tournament(st, p), ib = tournament(st, p);
    const double* pa = p.X + int64_t(ia) * p.d;
    const double* pb = p.X + int64_t(ib) * p.d;
    const bool do_pair = curand_uniform_double(&st) < p.pc;
    double* c1 = p.Xoff + int64_t(2 * pair) * p.d;
    double* c2 = (2 * pair + 1 < p.P) ? c1 + p.d : nullptr;
    const double pm = 1.0 / p.d;
    for (int k = 0; k < p.d; ++k) {
        const double lo = p.xl[k], hi = p.xu[k];
        double x1 = pa[k], x2 = pb[k];
        const double uv = curand_uniform_double(&st), u = curand_uniform_double(&st), us = curand_uniform_double(&st);
        if (do_pair && uv < 0.5 && fabs(x1 - x2) > 1e-14) {
            const double y1 = fmin(x1, x2), y2 = fmax(x1, x2), delta = y2 - y1;
            const double b1 = sbx_betaq(u, 1.0 + 2.0 * (y1 - lo) / delta, p.eta_c);
            const double b2 = sbx_betaq(u, 1.0 + 2.0 * (hi - y2) / delta, p.eta_c);
            const double o1 = fmin(fmax(0.5 * ((y1 + y2) - b1 * delta), lo), hi);
            const double o2 = fmin(fmax(0.5 * ((y1 + y2) + b2 * delta), lo), hi);
            x1 = us < 0.5 ? o2 : o1;
            x2 = us < 0.5 ? o1 : o2;
        }
        const double m1 = curand_uniform_double(&st), v1 = curand_uniform_double(&st);
        const double m2 = curand_uniform_double(&st), v2 = curand_uniform_double(&st);
        if (m1 < pm) x1 = poly_mutate(x1, lo, hi, v1, p.eta_m);
        if (m2 < pm) x2 = poly_mutate(x2, lo, hi, v2, p.eta_m);
        c1[k] = x1;
        if (c2) c2[k] = x2;
    }
}

// Xn = clip((X - xl) / (xu - xl), 0, 1): the normalisation SurrogateModel.evaluate applies (utils/normalization.py:31-32)
__global__ void normalize_kernel(const double* __restrict__ X, const double* __restrict__ xl, const double* __restrict__ xu,
                                 int64_t n, int d, double* __restrict__ Xn) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n * d) return;
    const int k = int(i % d);
    Xn[i] = fmin(fmax((X[i] - xl[k]) / (xu[k] - xl[k]), 0.0), 1.0);
}

// dup[o] = 1 when offspring o equals (on the 1e-8 grid) a population member or an EARLIER offspring
__global__ void __launch_bounds__(128) duplicate_kernel(const double* __restrict__ Xpop, int n_cur, const double* __restrict__ Xoff,
                                                        int P, int d, uint8_t* __restrict__ dup) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= P) return;
    double q[NS_MAX_D];
    for (int k = 0; k < d; ++k) q[k] = rint(Xoff[int64_t(o) * d + k] * 1e8);
    bool found = false;
    for (int j = 0; j < n_cur + o && !found; ++j) {
        const double* x = j < n_cur ? Xpop + int64_t(j) * d : Xoff + int64_t(j - n_cur) * d;
        bool same = true;
        for (int k = 0; k < d && same; ++k) same = rint(x[k] * 1e8) == q[k];
        found = same;
    }
    dup[o] = found ? 1 : 0;
}

// merged objective-major table YT[k][i] of population (rows < n_cur) and offspring; discarded offspring get +DBL_MAX
__global__ void merge_kernel(const double* __restrict__ Fpop, int n_cur, const double* __restrict__ Foff, const uint8_t* __restrict__ dup,
                             int P, int m, double* __restrict__ YT) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n = n_cur + P;
    if (i >= n) return;
    for (int k = 0; k < m; ++k) {
        double v;
        if (i < n_cur) v = Fpop[int64_t(i) * m + k];
        else v = (dup && dup[i - n_cur]) ? DBL_MAX : Foff[int64_t(i - n_cur) * m + k];
        YT[int64_t(k) * n + i] = v;
    }
}

// ---- crowding distance + survival in one CTA (n <= 4096) --------------------------------------------------------
template <class Less>
__device__ void bitonic_sort(int* idx, int np2, Less less) {
    for (int k = 2; k <= np2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < np2; t += NS_SORT_THREADS) {
                const int u = t ^ j;
                if (u > t) {
                    const int a = idx[t], b = idx[u];
                    const bool up = (t & k) == 0;
                    if (less(b, a) == up) {
                        idx[t] = b;
                        idx[u] = a;
                    }
                }
            }
            __syncthreads();
        }
}

struct SurviveArgs {
    const double* YT;   // [m][n]
    const int* rank;    // [n]
    double* crowd;      // [n] scratch / output (crowding distance of every merged individual)
    int* sel;           // [keep] survivors (indices into the merged table), best first
    int n, m, keep;
};

__global__ void __launch_bounds__(NS_SORT_THREADS) survive_kernel(SurviveArgs p) {
    extern __shared__ int sv_smem[];
    const int n = p.n;
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    int* idx = sv_smem;          // [np2]
    int* gstart = idx + np2;     // [n] first sorted position of each rank group
    int* gend = gstart + n;      // [n] last sorted position of each rank group
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += NS_SORT_THREADS) p.crowd[i] = 0.0;
    for (int k = 0; k < p.m; ++k) {
        const double* f = p.YT + int64_t(k) * n;
        for (int t = tid; t < np2; t += NS_SORT_THREADS) idx[t] = t;  // indices >= n are padding (sorted last)
        __syncthreads();
        bitonic_sort(idx, np2, [&](int a, int b) {
            if (a >= n || b >= n) return a < n && b >= n ? true : (a >= n && b >= n ? a < b : false);
            const int ra = p.rank[a], rb = p.rank[b];
            if (ra != rb) return ra < rb;
            if (f[a] != f[b]) return f[a] < f[b];
            return a < b;
        });
        for (int t = tid; t < n; t += NS_SORT_THREADS) {
            const int r = p.rank[idx[t]];
            if (t == 0 || p.rank[idx[t - 1]] != r) gstart[r] = t;
            if (t == n - 1 || p.rank[idx[t + 1]] != r) gend[r] = t;
        }
        __syncthreads();
        for (int t = tid; t < n; t += NS_SORT_THREADS) {
            const int i = idx[t], r = p.rank[i];
            const int s = gstart[r], e = gend[r];
            if (t == s || t == e) {
                p.crowd[i] = INFINITY;  // boundary points of a front
            } else {
                const double span = f[idx[e]] - f[idx[s]];
                if (span > 0.0 && span < DBL_MAX) p.crowd[i] += (f[idx[t + 1]] - f[idx[t - 1]]) / span;
            }
        }
        __syncthreads();
    }
    // survival: fronts in order, inside a front the lonelier individual first
    for (int t = tid; t < np2; t += NS_SORT_THREADS) idx[t] = t;
    __syncthreads();
    bitonic_sort(idx, np2, [&](int a, int b) {
        if (a >= n || b >= n) return a < n && b >= n ? true : (a >= n && b >= n ? a < b : false);
        const int ra = p.rank[a], rb = p.rank[b];
        if (ra != rb) return ra < rb;
        const double ca = p.crowd[a], cb = p.crowd[b];
        if (ca != cb) return ca > cb;
        return a < b;
    });
    for (int t = tid; t < p.keep; t += NS_SORT_THREADS) p.sel[t] = idx[t];
}

// new population = merged[sel]
__global__ void gather_kernel(const double* __restrict__ Xpop, const double* __restrict__ Fpop, int n_cur, const double* __restrict__ Xoff,
                              const double* __restrict__ Foff, const int* __restrict__ rank, const double* __restrict__ crowd,
                              const int* __restrict__ sel, int keep, int d, int m, double* __restrict__ Xnew, double* __restrict__ Fnew,
                              int* __restrict__ rnew, double* __restrict__ cnew) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= keep) return;
    const int i = sel[t];
    const double* x = i < n_cur ? Xpop + int64_t(i) * d : Xoff + int64_t(i - n_cur) * d;
    const double* f = i < n_cur ? Fpop + int64_t(i) * m : Foff + int64_t(i - n_cur) * m;
    for (int k = 0; k < d; ++k) Xnew[int64_t(t) * d + k] = x[k];
    for (int k = 0; k < m; ++k) Fnew[int64_t(t) * m + k] = f[k];
    rnew[t] = rank[i];
    cnew[t] = crowd[i];
}

__global__ void count_valid_kernel(const uint8_t* dup, int P, unsigned long long* n_eval) {
    __shared__ unsigned int c;
    if (threadIdx.x == 0) c = 0;
    __syncthreads();
    unsigned int mine = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) mine += dup[i] ? 0 : 1;
    atomicAdd(&c, mine);
    __syncthreads();
    if (threadIdx.x == 0) *n_eval += c;
}

struct NBuf {
    void* p = nullptr;
    ~NBuf() {
        if (p) cudaFree(p);
    }
    template <typename T>
    T* as() { return static_cast<T*>(p); }
};

template <int M>
int launch_rank_m(const double* YT, int n, int32_t* rank, cudaStream_t s) {
    const size_t sm = rank_small_smem(M, n);
    NCU(cudaFuncSetAttribute(rank_small_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm)));
    rank_small_kernel<M><<<1, RS_THREADS, sm, s>>>(YT, n, rank);
    return AOED_OK;
}

int launch_rank(int m, const double* YT, int n, int32_t* rank, cudaStream_t s) {
    switch (m) {
        case 1: return launch_rank_m<1>(YT, n, rank, s);
        case 2: return launch_rank_m<2>(YT, n, rank, s);
        case 3: return launch_rank_m<3>(YT, n, rank, s);
        case 4: return launch_rank_m<4>(YT, n, rank, s);
        case 5: return launch_rank_m<5>(YT, n, rank, s);
        case 6: return launch_rank_m<6>(YT, n, rank, s);
        case 7: return launch_rank_m<7>(YT, n, rank, s);
        case 8: return launch_rank_m<8>(YT, n, rank, s);
        default: return nfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    }
}

}  // namespace

namespace {

// evaluator: (stream, normalised candidates [rows][d] on the device, rows, output [rows][m] on the device) -> rc
using Evaluator = std::function<int(cudaStream_t, const double*, int64_t, double*)>;

int nsga2_core(int device, int n_var, int n_obj, const Evaluator& eval_fn, int64_t n_init, const double* h_X0, const double* h_xl,
               const double* h_xu, int pop_size, int n_gen, uint64_t seed, double* h_X, double* h_F, int64_t* h_n_out,
               int64_t* h_n_eval) {
    if (!h_X0 || !h_xl || !h_xu || !h_X || !h_F || !h_n_out) return nfail(AOED_ERR_INVALID, "null argument");
    if (n_init < 1 || pop_size < 2 || n_gen < 1) return nfail(AOED_ERR_INVALID, "n_init >= 1, pop_size >= 2, n_gen >= 1 required");
    if (n_var < 1 || n_var > NS_MAX_D) return nfail(AOED_ERR_UNSUPPORTED, "n_var > 32 is not supported yet");
    const int d = n_var, m = n_obj, P = pop_size;
    const int64_t cap = std::max<int64_t>(n_init, P);  // the initial sampling is taken as is, even when larger than pop_size
    if (cap + P > RS_MAX_N || rank_small_smem(m, int(cap + P)) > 200 * 1024)
        return nfail(AOED_ERR_UNSUPPORTED, "population + offspring must fit the one-CTA sort (<= 4096 individuals)");
    DeviceGuard guard_(device);
    NCU(guard_.err);
    cudaStream_t s = nullptr;
    NCU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamDestroy(s); }
    } sguard{s};
    NBuf bX[2], bF[2], bR[2], bC[2], bXoff, bXn, bFoff, bDup, bYT, bRank, bCrowd, bSel, bXl, bXu, bEval;
    for (int i = 0; i < 2; ++i) {
        NCU(cudaMalloc(&bX[i].p, size_t(cap) * d * sizeof(double)));
        NCU(cudaMalloc(&bF[i].p, size_t(cap) * m * sizeof(double)));
        NCU(cudaMalloc(&bR[i].p, size_t(cap) * sizeof(int)));
        NCU(cudaMalloc(&bC[i].p, size_t(cap) * sizeof(double)));
    }
    NCU(cudaMalloc(&bXoff.p, size_t(P) * d * sizeof(double)));
    NCU(cudaMalloc(&bXn.p, size_t(cap) * d * sizeof(double)));
    NCU(cudaMalloc(&bFoff.p, size_t(P) * m * sizeof(double)));
    NCU(cudaMalloc(&bDup.p, size_t(P)));
    NCU(cudaMalloc(&bYT.p, size_t(cap + P) * m * sizeof(double)));
    NCU(cudaMalloc(&bRank.p, size_t(cap + P) * sizeof(int)));
    NCU(cudaMalloc(&bCrowd.p, size_t(cap + P) * sizeof(double)));
    NCU(cudaMalloc(&bSel.p, size_t(cap) * sizeof(int)));
    NCU(cudaMalloc(&bXl.p, d * sizeof(double)));
    NCU(cudaMalloc(&bXu.p, d * sizeof(double)));
    NCU(cudaMalloc(&bEval.p, sizeof(unsigned long long)));
    NCU(cudaMemcpyAsync(bXl.p, h_xl, d * sizeof(double), cudaMemcpyHostToDevice, s));
    NCU(cudaMemcpyAsync(bXu.p, h_xu, d * sizeof(double), cudaMemcpyHostToDevice, s));
    NCU(cudaMemcpyAsync(bX[0].p, h_X0, size_t(n_init) * d * sizeof(double), cudaMemcpyHostToDevice, s));
    NCU(cudaMemsetAsync(bEval.p, 0, sizeof(unsigned long long), s));

    auto evaluate = [&](const double* dXc, int64_t rows, double* dF) -> int {
        normalize_kernel<<<unsigned((rows * d + 255) / 256), 256, 0, s>>>(dXc, bXl.as<double>(), bXu.as<double>(), rows, d, bXn.as<double>());
        return eval_fn(s, bXn.as<double>(), rows, dF);
    };
    // rank + crowding of `n` individuals whose objectives are Fa (first na rows) and Fb; keeps the best `keep` in buffer `dst`
    auto survive = [&](int src, int na, const double* Xb, const double* Fb, const uint8_t* dup, int nb, int keep, int dst) -> int {
        const int n = na + nb;
        merge_kernel<<<(n + 255) / 256, 256, 0, s>>>(bF[src].as<double>(), na, Fb, dup, nb, m, bYT.as<double>());
        NCHK(launch_rank(m, bYT.as<double>(), n, bRank.as<int>(), s));
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        SurviveArgs sa;
        sa.YT = bYT.as<double>(); sa.rank = bRank.as<int>(); sa.crowd = bCrowd.as<double>(); sa.sel = bSel.as<int>();
        sa.n = n; sa.m = m; sa.keep = keep;
        const size_t sm = (size_t(np2) + 2 * size_t(n)) * sizeof(int);
        survive_kernel<<<1, NS_SORT_THREADS, sm, s>>>(sa);
        gather_kernel<<<(keep + 127) / 128, 128, 0, s>>>(bX[src].as<double>(), bF[src].as<double>(), na, Xb, Fb, bRank.as<int>(),
                                                          bCrowd.as<double>(), bSel.as<int>(), keep, d, m, bX[dst].as<double>(),
                                                          bF[dst].as<double>(), bR[dst].as<int>(), bC[dst].as<double>());
        NCU(cudaGetLastError());
        return AOED_OK;
    };

    // generation 1 = evaluation of the given sampling, truncated only when it exceeds pop_size
    NCHK(evaluate(bX[0].as<double>(), n_init, bF[0].as<double>()));
    int n_cur = int(std::min<int64_t>(n_init, P));
    NCHK(survive(0, int(n_init), nullptr, nullptr, nullptr, 0, n_cur, 1));
    int cur = 1;
    int64_t evals = n_init;
    for (int g = 1; g < n_gen; ++g) {
        VarArgs va;
        va.X = bX[cur].as<double>(); va.rank = bR[cur].as<int>(); va.crowd = bC[cur].as<double>();
        va.xl = bXl.as<double>(); va.xu = bXu.as<double>(); va.Xoff = bXoff.as<double>();
        va.n_cur = n_cur; va.P = P; va.d = d; va.seed = seed; va.gen = g; va.pc = 0.9; va.eta_c = 15.0; va.eta_m = 20.0;
        variation_kernel<<<((P + 1) / 2 + 127) / 128, 128, 0, s>>>(va);
        duplicate_kernel<<<(P + 127) / 128, 128, 0, s>>>(bX[cur].as<double>(), n_cur, bXoff.as<double>(), P, d, bDup.as<uint8_t>());
        count_valid_kernel<<<1, 256, 0, s>>>(bDup.as<uint8_t>(), P, bEval.as<unsigned long long>());
        NCHK(evaluate(bXoff.as<double>(), P, bFoff.as<double>()));
        NCHK(survive(cur, n_cur, bXoff.as<double>(), bFoff.as<double>(), bDup.as<uint8_t>(), P, P, cur ^ 1));
        cur ^= 1;
        n_cur = P;
    }
    NCU(cudaMemcpyAsync(h_X, bX[cur].p, size_t(n_cur) * d * sizeof(double), cudaMemcpyDeviceToHost, s));
    NCU(cudaMemcpyAsync(h_F, bF[cur].p, size_t(n_cur) * m * sizeof(double), cudaMemcpyDeviceToHost, s));
    unsigned long long off_evals = 0;
    NCU(cudaMemcpyAsync(&off_evals, bEval.p, sizeof(off_evals), cudaMemcpyDeviceToHost, s));
    NCU(cudaStreamSynchronize(s));
    *h_n_out = n_cur;
    if (h_n_eval) *h_n_eval = evals + int64_t(off_evals);
    return AOED_OK;
}

}  // namespace

extern "C" int aoed_nsga2_run(aoed_gp_t* gp, int device, int n_var, int n_obj, int acq_kind, const double* h_acq_param,
                              int64_t n_init, const double* h_X0, const double* h_xl, const double* h_xu, int pop_size,
                              int n_gen, uint64_t seed, double* h_X, double* h_F, int64_t* h_n_out, int64_t* h_n_eval) {
    if (!gp) return nfail(AOED_ERR_INVALID, "null handle");
    if (acq_kind < AOED_ACQ_IDENTITY || acq_kind > AOED_ACQ_PI) return nfail(AOED_ERR_INVALID, "an acquisition kind is required");
    const uint32_t flags = acq_kind >= AOED_ACQ_UCB ? AOED_EVAL_STD : 0u;
    Evaluator ev = [=](cudaStream_t s, const double* dXn, int64_t rows, double* dF) -> int {
        return aoed_gp_eval_device(gp, rows, dXn, flags, acq_kind, h_acq_param, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                   dF, nullptr, nullptr, s);
    };
    return nsga2_core(device, n_var, n_obj, ev, n_init, h_X0, h_xl, h_xu, pop_size, n_gen, seed, h_X, h_F, h_n_out, h_n_eval);
}

extern "C" int aoed_nsga2_run_rff(int device, int n_var, int n_obj, int n_feat, const double* h_W, const double* h_b,
                                  const double* h_theta, const double* h_factor, const double* h_y_mean, const double* h_y_scale,
                                  int64_t n_init, const double* h_X0, const double* h_xl, const double* h_xu, int pop_size,
                                  int n_gen, uint64_t seed, double* h_X, double* h_F, int64_t* h_n_out, int64_t* h_n_eval) {
    if (!h_W || !h_b || !h_theta || !h_factor || !h_y_mean || !h_y_scale || n_feat < 1 || n_var < 1 || n_var > NS_MAX_D)
        return nfail(AOED_ERR_INVALID, "bad argument");
    if (rff_smem_bytes(n_feat, rff_dp(n_var)) > 200 * 1024) return nfail(AOED_ERR_UNSUPPORTED, "too many spectral points for shared memory");
    DeviceGuard guard_(device);
    NCU(guard_.err);
    const size_t d = n_var, m = n_obj, M = n_feat;
    NBuf par;
    NCU(cudaMalloc(&par.p, (m * M * d + 2 * m * M + 3 * m) * sizeof(double)));
    double* P = par.as<double>();
    RffArgs base;
    base.W = P; base.b = P + m * M * d; base.theta = base.b + m * M; base.factor = base.theta + m * M;
    base.y_mean = base.factor + m; base.y_scale = base.y_mean + m;
    NCU(cudaMemcpy(const_cast<double*>(base.W), h_W, m * M * d * sizeof(double), cudaMemcpyHostToDevice));
    NCU(cudaMemcpy(const_cast<double*>(base.b), h_b, m * M * sizeof(double), cudaMemcpyHostToDevice));
    NCU(cudaMemcpy(const_cast<double*>(base.theta), h_theta, m * M * sizeof(double), cudaMemcpyHostToDevice));
    NCU(cudaMemcpy(const_cast<double*>(base.factor), h_factor, m * sizeof(double), cudaMemcpyHostToDevice));
    NCU(cudaMemcpy(const_cast<double*>(base.y_mean), h_y_mean, m * sizeof(double), cudaMemcpyHostToDevice));
    NCU(cudaMemcpy(const_cast<double*>(base.y_scale), h_y_scale, m * sizeof(double), cudaMemcpyHostToDevice));
    base.dF = base.hF = nullptr;
    base.d = n_var; base.m = n_obj; base.M = n_feat;
    Evaluator ev = [=](cudaStream_t s, const double* dXn, int64_t rows, double* dF) -> int {
        RffArgs a = base;
        a.X = dXn; a.F = dF; a.rows = rows;
        cudaError_t e = rff_launch_value(a, s);
        if (e != cudaSuccess) return nfail(AOED_ERR_CUDA, std::string("rff_kernel: ") + cudaGetErrorString(e));
        return AOED_OK;
    };
    return nsga2_core(device, n_var, n_obj, ev, n_init, h_X0, h_xl, h_xu, pop_size, n_gen, seed, h_X, h_F, h_n_out, h_n_eval);
}
