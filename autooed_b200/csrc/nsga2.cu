// Device-resident NSGA-II generation loop on the GP acquisition (sm_100a) — SURVEY.md §8 f1.
//
// Replaces, for the common unconstrained case, the pymoo 0.4.1 `NSGA2` + `minimize(..., ('n_gen', G))` loop that
// `autooed/mobo/solver/nsga2.py:17-30` drives (one `SurrogateProblem.evaluate` per generation): variation, duplicate
// elimination, acquisition evaluation, non-dominated sorting, crowding distance and survival all stay on the GPU; the
// host only enqueues ~12 launches per generation and never synchronises inside the loop.
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
#include <chrono>
#include <cstdio>
#include <cstdlib>
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

constexpr int NS_MAX_D = 64;

struct VarArgs {
    const double* X;      // [n_cur][d] current population (continuous)
    const double* F;      // [n_cur][m] its acquisition values
    const double* crowd;  // [n_cur]
    int m;
    const double *xl, *xu;
    double* Xoff;         // [P][d] offspring (continuous)
    double* Xn;           // [P][d] the same, normalised to the unit box as the surrogate wants them
    int n_cur, P, d;
    unsigned long long seed;
    const int* gen;       // generation number, on the device so that a captured graph of the loop can be replayed
    double pc, eta_c, eta_m;
};

// pymoo 0.4.1 NSGA2: binary tournament of type 'comp_by_dom_and_crowding' — the contestant that DOMINATES the other wins
// (Dominator.get_relation on the objective vectors), otherwise the larger crowding distance, a coin on ties
__device__ __forceinline__ int tournament(curandStatePhilox4_32_10_t& st, const VarArgs& p) {
    const int a = min(int(curand_uniform_double(&st) * p.n_cur), p.n_cur - 1);
    const int b = min(int(curand_uniform_double(&st) * p.n_cur), p.n_cur - 1);
    const double coin = curand_uniform_double(&st);
    bool a_lt = false, b_lt = false;
    for (int t = 0; t < p.m; ++t) {
        const double fa = p.F[int64_t(a) * p.m + t], fb = p.F[int64_t(b) * p.m + t];
        a_lt = a_lt || fa < fb;
        b_lt = b_lt || fb < fa;
    }
    if (a_lt != b_lt) return a_lt ? a : b;
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

// one thread per (offspring pair, variable).  The pair-level draws (two tournaments per parent, the crossover coin) sit at
// the start of the pair's Philox subsequence and are replayed by each of its threads; variable k owns the 64 draws at
// offset 64 (k + 1) of the same subsequence, so the result does not depend on how the work is spread over threads.
__global__ void __launch_bounds__(128) variation_kernel(VarArgs p) {
    pdl_wait();
    pdl_trigger();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = gid / p.d, k = gid - pair * p.d;
    if (2 * pair >= p.P) return;
    curandStatePhilox4_32_10_t st;
    const unsigned long long base = (unsigned long long)(*p.gen) * 8192ull;  // 64 (k + 1) <= 4160 draws per generation and pair
    curand_init(p.seed, (unsigned long long)pair, base, &st);
    const int ia = tournament(st, p), ib = tournament(st, p);
    const bool do_pair = curand_uniform_double(&st) < p.pc;
    curand_init(p.seed, (unsigned long long)pair, base + 64ull * (k + 1), &st);
    const double lo = p.xl[k], hi = p.xu[k];
    double x1 = p.X[int64_t(ia) * p.d + k], x2 = p.X[int64_t(ib) * p.d + k];
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
    const double pm = 1.0 / p.d;
    const double m1 = curand_uniform_double(&st), v1 = curand_uniform_double(&st);
    const double m2 = curand_uniform_double(&st), v2 = curand_uniform_double(&st);
    if (m1 < pm) x1 = poly_mutate(x1, lo, hi, v1, p.eta_m);
    if (m2 < pm) x2 = poly_mutate(x2, lo, hi, v2, p.eta_m);
    // also in normalised form, Xn = clip((X - xl) / (xu - xl), 0, 1) (utils/normalization.py:31-32; the expression of
    // normalize_kernel below), so that the evaluator starts right after this kernel
    p.Xoff[int64_t(2 * pair) * p.d + k] = x1;
    p.Xn[int64_t(2 * pair) * p.d + k] = fmin(fmax((x1 - lo) / (hi - lo), 0.0), 1.0);
    if (2 * pair + 1 < p.P) {
        p.Xoff[int64_t(2 * pair + 1) * p.d + k] = x2;
        p.Xn[int64_t(2 * pair + 1) * p.d + k] = fmin(fmax((x2 - lo) / (hi - lo), 0.0), 1.0);
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

// dup[o] = 1 when offspring o equals (on the 1e-8 grid) a population member or an EARLIER offspring.  The grid value of
// the first coordinate of every individual is staged in shared memory; one warp per offspring scans it and looks at the
// other coordinates only where the first one matches.  *n_eval += number of offspring kept; also advances *gen.
constexpr int DUP_WARPS = 8;
__global__ void __launch_bounds__(32 * DUP_WARPS) duplicate_kernel(const double* __restrict__ Xpop, int n_cur,
                                                                   const double* __restrict__ Xoff, int P, int d,
                                                                   uint8_t* __restrict__ dup, unsigned long long* n_eval,
                                                                   int* gen) {
    extern __shared__ double dup_key[];  // [n_cur + P] first coordinates: a cheap necessary condition before the full distance
    if (blockIdx.x == 0 && threadIdx.x == 0) *gen += 1;  // the variation kernel of this generation has read it
    for (int j = threadIdx.x; j < n_cur + P; j += 32 * DUP_WARPS)
        dup_key[j] = j < n_cur ? Xpop[int64_t(j) * d] : Xoff[int64_t(j - n_cur) * d];
    __syncthreads();
    const int o = blockIdx.x * DUP_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= P) return;
    const double* xo = Xoff + int64_t(o) * d;
    const double q0 = dup_key[n_cur + o];
    const int total = n_cur + o;
    bool same_any = false;
    // pymoo DefaultDuplicateElimination: Euclidean distance to a population member or an EARLIER offspring <= 1e-16
    for (int j = lane; j < total; j += 32) {
        if (fabs(dup_key[j] - q0) <= 1e-16) {  // rare: the full distance
            const double* x = j < n_cur ? Xpop + int64_t(j) * d : Xoff + int64_t(j - n_cur) * d;
            double d2 = 0.0;
            for (int k = 0; k < d; ++k) {
                const double t = x[k] - xo[k];
                d2 = fma(t, t, d2);
            }
            same_any = same_any || (sqrt(d2) <= 1e-16);
        }
    }
    const bool found = __any_sync(0xffffffffu, same_any);
    if (lane == 0) {
        dup[o] = found ? 1 : 0;
        if (!found) atomicAdd(n_eval, 1ull);
    }
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

// ---- crowding distance + survival by pairwise counting (no sort loop, many CTAs) ----------------------------
// A warp owns 32 consecutive individuals, the warps of a CTA split the individuals they are compared with, which are
// staged in shared memory first (see dom_count_kernel).  Individuals ranked above *last_rank (the unpeeled rest of an
// early-stopped sort) can never survive and are skipped ("relevant" = rank <= *last_rank).
constexpr int CW_WARPS = 16;
inline size_t crowd_smem(int n) { return size_t(n) * (sizeof(double) + 2 * sizeof(int)); }

struct SurvivalTables {
    const double* YT;     // [m][n] merged objectives
    const int* rank;      // [n]
    const int* last_rank;
    int* order;           // [m][n] order[k][t] = individual at position t when sorted by (rank, objective k, index)
    int* pos;             // [m][n] inverse of order (relevant individuals only)
    int* gbase;           // [n] first position of i's front
    int* gsize;           // [n] size of i's front
    int n, m;
};

// compacts the relevant individuals into shared memory: rs = rank, js = index; returns how many
__device__ int stage_relevant(const int* __restrict__ rank, int n, int last, int* rs, int* js, int* n_rel) {
    for (int j0 = 0; j0 < n; j0 += 32 * CW_WARPS) {
        const int j = j0 + threadIdx.x;
        const int r = j < n ? rank[j] : INT_MAX;
        const int q = warp_slot(n_rel, r <= last);
        if (r <= last) {
            rs[q] = r;
            js[q] = j;
        }
    }
    __syncthreads();
    return *n_rel;
}

// positions by counting: for every objective, where individual i stands when its front is ordered by (value, index) —
// the order a stable sort gives.  ~10 instructions per pair; the neighbour lookups happen in select_kernel.
__global__ void __launch_bounds__(32 * CW_WARPS) crowd_pos_kernel(SurvivalTables p) {
    extern __shared__ double cw_smem[];
    pdl_wait();
    pdl_trigger();
    const int n = p.n;
    double* fs = cw_smem;                      // [nr] objective k of the relevant individuals
    int* rs = reinterpret_cast<int*>(fs + n);  // [nr] ranks
    int* js = rs + n;                          // [nr] indices
    __shared__ int c_pos[32], c_base[32], c_size[32];
    __shared__ int n_rel;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = blockIdx.x * 32 + lane;
    const int last = *p.last_rank;
    const int ri = i < n ? p.rank[i] : INT_MAX;
    const bool active = i < n && ri <= last;
    if (threadIdx.x == 0) n_rel = 0;
    if (__syncthreads_or(active) == 0) return;
    const int nr = stage_relevant(p.rank, n, last, rs, js, &n_rel);
    int base = 0;
    for (int k = 0; k < p.m; ++k) {
        const double* f = p.YT + int64_t(k) * n;
        for (int q = threadIdx.x; q < nr; q += 32 * CW_WARPS) fs[q] = f[js[q]];
        if (w == 0) c_pos[lane] = c_base[lane] = c_size[lane] = 0;
        __syncthreads();
        if (active) {
            const double fi = f[i];
            int c = 0, cb = 0, cs = 0;
#pragma unroll 4
            for (int q = w; q < nr; q += CW_WARPS) {
                const int rj = rs[q];
                const double fj = fs[q];
                const bool mine = rj == ri;
                c += (mine && (fj < fi || (fj == fi && js[q] < i))) ? 1 : 0;
                cb += rj < ri ? 1 : 0;
                cs += mine ? 1 : 0;
            }
            atomicAdd(&c_pos[lane], c);
            if (k == 0) {
                atomicAdd(&c_base[lane], cb);
                atomicAdd(&c_size[lane], cs);
            }
        }
        __syncthreads();
        if (w == 0 && active) {
            if (k == 0) {
                base = c_base[lane];
                p.gbase[i] = base;
                p.gsize[i] = c_size[lane];
            }
            const int t = base + c_pos[lane];
            p.pos[int64_t(k) * n + i] = t;
            p.order[int64_t(k) * n + t] = i;
        }
        __syncthreads();
    }
}

// crowding distance of a relevant individual from the position tables (pymoo calc_crowding_distance: neighbours in its
// own front per objective, normalised by the front's extent; front boundary -> +inf)
__device__ __forceinline__ double crowding_of(const SurvivalTables& p, int i) {
    const int b = p.gbase[i], e = b + p.gsize[i] - 1;
    double acc = 0.0;
    for (int k = 0; k < p.m; ++k) {
        const int* o = p.order + int64_t(k) * p.n;
        const double* f = p.YT + int64_t(k) * p.n;
        const int t = p.pos[int64_t(k) * p.n + i];
        if (t == b || t == e) {
            acc = INFINITY;
        } else {
            const double span = f[o[e]] - f[o[b]];
            if (span > 0.0 && span < DBL_MAX) acc += (f[o[t + 1]] - f[o[t - 1]]) / span;
        }
    }
    return acc;
}

// survivors: position of i in the order (rank, crowding distance descending, index) = number of individuals before it;
// the individuals whose position is below `keep` are copied to that row of the next population
struct SelectArgs {
    SurvivalTables tab;
    double* crowd;           // [n] out: crowding distance (0 for the irrelevant rest)
    const double *Xpop, *Fpop, *Xoff, *Foff;  // merged table = population rows [0, n_cur) then offspring rows
    int n_cur, keep, d;
    int* sel;                // [keep] survivors' indices in the merged table, best first
    double *Xnew, *Fnew, *cnew;
    int* rnew;
};

__global__ void __launch_bounds__(32 * CW_WARPS) select_kernel(SelectArgs p) {
    extern __shared__ double cw_smem[];
    pdl_wait();
    pdl_trigger();
    const int n = p.tab.n;
    double* cs = cw_smem;                      // [nr] crowding distances of the relevant individuals
    int* rs = reinterpret_cast<int*>(cs + n);  // [nr] ranks
    int* js = rs + n;                          // [nr] indices
    __shared__ int part[32];
    __shared__ int n_rel;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = blockIdx.x * 32 + lane;
    const int last = *p.tab.last_rank;
    const int ri = i < n ? p.tab.rank[i] : INT_MAX;
    const bool active = i < n && ri <= last;
    if (threadIdx.x == 0) n_rel = 0;
    if (w == 0) part[lane] = 0;
    if (__syncthreads_or(active) == 0) {
        if (w == 0 && i < n) p.crowd[i] = 0.0;
        return;
    }
    const int nr = stage_relevant(p.tab.rank, n, last, rs, js, &n_rel);
    for (int q = threadIdx.x; q < nr; q += 32 * CW_WARPS) cs[q] = crowding_of(p.tab, js[q]);
    __syncthreads();
    const double ci = active ? crowding_of(p.tab, i) : 0.0;
    if (active) {
        int c = 0;
#pragma unroll 4
        for (int q = w; q < nr; q += CW_WARPS) {
            const int rj = rs[q], j = js[q];
            const double cj = cs[q];
            c += (rj != ri ? rj < ri : (cj != ci ? cj > ci : j < i)) ? 1 : 0;
        }
        atomicAdd(&part[lane], c);
    }
    __syncthreads();
    if (w == 0 && i < n) p.crowd[i] = ci;
    if (w == 0 && active && part[lane] < p.keep) {
        const int t = part[lane];
        p.sel[t] = i;
        if (p.Xnew) {
            const int m = p.tab.m;
            const double* x = i < p.n_cur ? p.Xpop + int64_t(i) * p.d : p.Xoff + int64_t(i - p.n_cur) * p.d;
            const double* f = i < p.n_cur ? p.Fpop + int64_t(i) * m : p.Foff + int64_t(i - p.n_cur) * m;
            for (int k = 0; k < p.d; ++k) p.Xnew[int64_t(t) * p.d + k] = x[k];
            for (int k = 0; k < m; ++k) p.Fnew[int64_t(t) * m + k] = f[k];
            p.rnew[t] = ri;
            p.cnew[t] = ci;
        }
    }
}

struct NBuf {
    void* p = nullptr;
    ~NBuf() {
        if (p) cudaFree(p);
    }
    template <typename T>
    T* as() { return static_cast<T*>(p); }
};

// the scratch tables of one survival step for up to n individuals and m objectives
struct SurvivalScratch {
    NBuf order, pos, gbase, gsize;
    int alloc(int n, int m) {
        NCU(cudaMalloc(&order.p, size_t(n) * m * sizeof(int)));
        NCU(cudaMalloc(&pos.p, size_t(n) * m * sizeof(int)));
        NCU(cudaMalloc(&gbase.p, size_t(n) * sizeof(int)));
        NCU(cudaMalloc(&gsize.p, size_t(n) * sizeof(int)));
        return AOED_OK;
    }
    SurvivalTables tables(const double* YT, const int* rank, const int* last_rank, int n, int m) {
        SurvivalTables t;
        t.YT = YT; t.rank = rank; t.last_rank = last_rank; t.n = n; t.m = m;
        t.order = order.as<int>(); t.pos = pos.as<int>(); t.gbase = gbase.as<int>(); t.gsize = gsize.as<int>();
        return t;
    }
};

int launch_rank(int m, const double* YT, int n, int32_t* rank, int* cnt_scratch, int need, int* last_rank, cudaStream_t s,
                bool pdl = false) {
#define NS_RANK_CASE(MM)                                                                   \
    case MM:                                                                               \
        NCU(launch_rank_small_m<MM>(YT, n, rank, cnt_scratch, need, s, last_rank, pdl));   \
        break;
    switch (m) {
        NS_RANK_CASE(1) NS_RANK_CASE(2) NS_RANK_CASE(3) NS_RANK_CASE(4) NS_RANK_CASE(5) NS_RANK_CASE(6) NS_RANK_CASE(7) NS_RANK_CASE(8)
        default: return nfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    }
#undef NS_RANK_CASE
    return AOED_OK;
}

}  // namespace

namespace {

// opt-in to the shared memory the staging kernels need for n individuals (above the 48 KB default from n ~ 4000)
int set_survival_smem(int n) {
    NCU(cudaFuncSetAttribute(crowd_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(crowd_smem(n))));
    NCU(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(crowd_smem(n))));
    return AOED_OK;
}

// evaluator: (stream, normalised candidates [rows][d] on the device, rows, output [rows][m] on the device) -> rc
using Evaluator = std::function<int(cudaStream_t, const double*, int64_t, double*)>;

int nsga2_core(int device, int n_var, int n_obj, const Evaluator& eval_fn, int64_t n_init, const double* h_X0, const double* h_xl,
               const double* h_xu, int pop_size, int n_gen, uint64_t seed, double* h_X, double* h_F, int64_t* h_n_out,
               int64_t* h_n_eval) {
    if (!h_X0 || !h_xl || !h_xu || !h_X || !h_F || !h_n_out) return nfail(AOED_ERR_INVALID, "null argument");
    if (n_init < 1 || pop_size < 2 || n_gen < 1) return nfail(AOED_ERR_INVALID, "n_init >= 1, pop_size >= 2, n_gen >= 1 required");
    if (n_var < 1 || n_var > NS_MAX_D) return nfail(AOED_ERR_UNSUPPORTED, "n_var > 64 is not supported by the device-resident loop");
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
    NBuf bX[2], bF[2], bR[2], bC[2], bXoff, bXn, bFoff, bDup, bYT, bRank, bCrowd, bSel, bXl, bXu, bEval, bCnt, bLast, bGen;
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
    NCU(cudaMalloc(&bCnt.p, rank_scratch_ints(int(cap + P)) * sizeof(int)));
    NCU(cudaMalloc(&bSel.p, size_t(cap) * sizeof(int)));
    NCU(cudaMalloc(&bXl.p, d * sizeof(double)));
    NCU(cudaMalloc(&bXu.p, d * sizeof(double)));
    NCU(cudaMalloc(&bEval.p, sizeof(unsigned long long)));
    NCU(cudaMemcpyAsync(bXl.p, h_xl, d * sizeof(double), cudaMemcpyHostToDevice, s));
    NCU(cudaMemcpyAsync(bXu.p, h_xu, d * sizeof(double), cudaMemcpyHostToDevice, s));
    NCU(cudaMemcpyAsync(bX[0].p, h_X0, size_t(n_init) * d * sizeof(double), cudaMemcpyHostToDevice, s));
    NCU(cudaMemsetAsync(bEval.p, 0, sizeof(unsigned long long), s));
    NCU(cudaMalloc(&bLast.p, sizeof(int)));
    NCU(cudaMalloc(&bGen.p, sizeof(int)));
    NCHK(set_survival_smem(int(cap + P)));
    SurvivalScratch scratch;
    NCHK(scratch.alloc(int(cap + P), m));

    // AOED_NSGA2_STAMP=1: CUDA events between the phases of every generation, per-phase totals on stderr (diagnostics)
    const bool stamp = getenv("AOED_NSGA2_STAMP") != nullptr;
    // AOED_NSGA2_PDL=0: ordinary stream serialisation between the kernels of a generation instead of programmatic dependent launch
    const char* pdl_env = getenv("AOED_NSGA2_PDL");
    const bool pdl = !stamp && !(pdl_env && pdl_env[0] == '0');
    std::vector<std::pair<int, cudaEvent_t>> marks;
    auto mark = [&](int phase) {
        if (!stamp) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        marks.emplace_back(phase, e);
    };

    auto evaluate = [&](const double* dXc, int64_t rows, double* dF) -> int {
        normalize_kernel<<<unsigned((rows * d + 255) / 256), 256, 0, s>>>(dXc, bXl.as<double>(), bXu.as<double>(), rows, d, bXn.as<double>());
        return eval_fn(s, bXn.as<double>(), rows, dF);
    };
    // rank + crowding of `n` individuals whose objectives are Fa (first na rows) and Fb; keeps the best `keep` in buffer `dst`
    auto survive = [&](int src, int na, const double* Xb, const double* Fb, const uint8_t* dup, int nb, int keep, int dst) -> int {
        const int n = na + nb;
        merge_kernel<<<(n + 255) / 256, 256, 0, s>>>(bF[src].as<double>(), na, Fb, dup, nb, m, bYT.as<double>());
        NCHK(launch_rank(m, bYT.as<double>(), n, bRank.as<int>(), bCnt.as<int>(), keep, bLast.as<int>(), s, pdl));
        mark(4);
        SelectArgs sa;
        sa.tab = scratch.tables(bYT.as<double>(), bRank.as<int>(), bLast.as<int>(), n, m);
        NCU(launch_pdl(pdl, crowd_pos_kernel, dim3((n + 31) / 32), dim3(32 * CW_WARPS), crowd_smem(n), s, sa.tab));
        mark(5);
        sa.crowd = bCrowd.as<double>();
        sa.Xpop = bX[src].as<double>(); sa.Fpop = bF[src].as<double>(); sa.Xoff = Xb; sa.Foff = Fb;
        sa.n_cur = na; sa.keep = keep; sa.d = d; sa.sel = bSel.as<int>();
        sa.Xnew = bX[dst].as<double>(); sa.Fnew = bF[dst].as<double>(); sa.cnew = bC[dst].as<double>(); sa.rnew = bR[dst].as<int>();
        NCU(launch_pdl(pdl, select_kernel, dim3((n + 31) / 32), dim3(32 * CW_WARPS), crowd_smem(n), s, sa));
        mark(6);
        NCU(cudaGetLastError());
        return AOED_OK;
    };

    // generation 1 = evaluation of the given sampling.  pymoo's `_initialize` runs the survival with
    // n_survive = len(pop): the WHOLE sampling (ranked, with crowding distances) parents the first offspring, the cut to
    // pop_size happens at the first merge
    NCHK(evaluate(bX[0].as<double>(), n_init, bF[0].as<double>()));
    int n_cur = int(n_init);
    NCHK(survive(0, int(n_init), nullptr, nullptr, nullptr, 0, n_cur, 1));
    int cur = 1;
    int64_t evals = n_init;
    const auto t_loop = std::chrono::steady_clock::now();
    const int one = 1;
    NCU(cudaMemcpyAsync(bGen.p, &one, sizeof(int), cudaMemcpyHostToDevice, s));
    // The duplicate scan needs only the offspring coordinates, the evaluator only their normalised copy: the two run side by
    // side (second stream; in the captured graph two branches that join at the merge), which takes the scan — 6 us at
    // P = 200, 20 us at P = 1000 — off the critical path of a generation.
    cudaStream_t s2 = nullptr;
    NCU(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    StreamGuard sguard2{s2};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    NCU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    NCU(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    struct EventGuard {
        cudaEvent_t a, b;
        ~EventGuard() {
            cudaEventDestroy(a);
            cudaEventDestroy(b);
        }
    } eguard{ev_fork, ev_join};
    const char* fork_env = getenv("AOED_NSGA2_FORK");  // =0: one serial stream (the per-phase stamps want that too)
    const bool serial = stamp || (fork_env && fork_env[0] == '0');
    auto generation = [&]() -> int {
        VarArgs va;
        va.X = bX[cur].as<double>(); va.F = bF[cur].as<double>(); va.m = m; va.crowd = bC[cur].as<double>();
        va.xl = bXl.as<double>(); va.xu = bXu.as<double>(); va.Xoff = bXoff.as<double>(); va.Xn = bXn.as<double>();
        va.n_cur = n_cur; va.P = P; va.d = d; va.seed = seed; va.gen = bGen.as<int>(); va.pc = 0.9; va.eta_c = 15.0; va.eta_m = 20.0;
        mark(0);
        NCU(launch_pdl(pdl, variation_kernel, dim3(((P + 1) / 2 * d + 127) / 128), dim3(128), 0, s, va));
        mark(1);
        cudaStream_t sdup = serial ? s : s2;
        if (!serial) {
            NCU(cudaEventRecord(ev_fork, s));
            NCU(cudaStreamWaitEvent(s2, ev_fork, 0));
        }
        duplicate_kernel<<<(P + DUP_WARPS - 1) / DUP_WARPS, 32 * DUP_WARPS, size_t(n_cur + P) * sizeof(double), sdup>>>(
            bX[cur].as<double>(), n_cur, bXoff.as<double>(), P, d, bDup.as<uint8_t>(), bEval.as<unsigned long long>(), bGen.as<int>());
        if (!serial) NCU(cudaEventRecord(ev_join, s2));
        mark(2);
        NCHK(eval_fn(s, bXn.as<double>(), P, bFoff.as<double>()));
        if (!serial) NCU(cudaStreamWaitEvent(s, ev_join, 0));
        mark(3);
        NCHK(survive(cur, n_cur, bXoff.as<double>(), bFoff.as<double>(), bDup.as<uint8_t>(), P, P, cur ^ 1));
        cur ^= 1;
        n_cur = P;
        return AOED_OK;
    };
    int g = 1;
    if (g < n_gen) {  // first offspring generation: population size settles at P, every workspace reaches its final size
        NCHK(generation());
        ++g;
    }
    // The rest of the loop has fixed shapes: capture two generations (the population buffers ping-pong) into a CUDA graph
    // and replay it — ~1 us between kernels instead of ~2.5 us for 11 stream launches per generation.
    // AOED_NSGA2_GRAPH=0 keeps plain stream launches.
    const char* graph_env = getenv("AOED_NSGA2_GRAPH");
    if (!stamp && !(graph_env && graph_env[0] == '0') && n_gen - g >= 6) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const int cur0 = cur;
        bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const int rc1 = generation(), rc2 = rc1 == AOED_OK ? generation() : rc1;
            ok = cudaStreamEndCapture(s, &graph) == cudaSuccess && rc2 == AOED_OK && graph != nullptr;
        }
        ok = ok && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        cur = cur0;  // nothing ran during capture
        if (ok) {
            for (; g + 1 < n_gen && ok; g += 2) ok = cudaGraphLaunch(exec, s) == cudaSuccess;
            if (!ok) {
                if (exec) cudaGraphExecDestroy(exec);
                if (graph) cudaGraphDestroy(graph);
                return nfail(AOED_ERR_CUDA, "cudaGraphLaunch failed in the NSGA-II loop");
            }
        } else {
            cudaGetLastError();  // capture refused (e.g. an evaluator that synchronises): plain launches below
        }
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
    }
    for (; g < n_gen; ++g) NCHK(generation());
    const auto t_enqueued = std::chrono::steady_clock::now();
    n_cur = std::min(n_cur, P);  // n_gen == 1 with a sampling larger than pop_size: the best pop_size rows (the output capacity)
    NCU(cudaMemcpyAsync(h_X, bX[cur].p, size_t(n_cur) * d * sizeof(double), cudaMemcpyDeviceToHost, s));
    NCU(cudaMemcpyAsync(h_F, bF[cur].p, size_t(n_cur) * m * sizeof(double), cudaMemcpyDeviceToHost, s));
    unsigned long long off_evals = 0;
    NCU(cudaMemcpyAsync(&off_evals, bEval.p, sizeof(off_evals), cudaMemcpyDeviceToHost, s));
    NCU(cudaStreamSynchronize(s));
    if (stamp) {
        fprintf(stderr, "[nsga2 stamp] host enqueue %8.1f us/generation, until drained %8.1f us/generation\n",
                std::chrono::duration<double, std::micro>(t_enqueued - t_loop).count() / std::max(1, n_gen - 1),
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_loop).count() / std::max(1, n_gen - 1));
        static const char* names[8] = {"", "variation", "duplicates", "evaluate", "merge+rank", "crowding", "select", ""};
        double tot[8] = {0};
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0.f;
            if (marks[i].first > 0 && cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second) == cudaSuccess)
                tot[marks[i].first] += ms;
        }
        for (int ph = 1; ph < 7; ++ph) fprintf(stderr, "[nsga2 stamp] %-10s %8.1f us/generation\n", names[ph], tot[ph] * 1e3 / std::max(1, n_gen - 1));
        for (auto& mk : marks) cudaEventDestroy(mk.second);
    }
    *h_n_out = n_cur;
    if (h_n_eval) *h_n_eval = evals + int64_t(off_evals);
    return AOED_OK;
}

}  // namespace

extern "C" int aoed_nsga2_survival(int device, int64_t n64, int n_obj, const double* h_F, int64_t keep64, int32_t* h_rank,
                                   double* h_crowd, int32_t* h_sel) {
    if (n64 < 1 || keep64 < 1 || keep64 > n64 || n_obj < 1 || !h_F || !h_rank || !h_crowd || !h_sel)
        return nfail(AOED_ERR_INVALID, "bad argument");
    if (n64 > RS_MAX_N || n_obj > 8 || rank_small_smem(n_obj, int(n64)) > 200 * 1024)
        return nfail(AOED_ERR_UNSUPPORTED, "at most 4096 individuals and 8 objectives");
    const int n = int(n64), m = n_obj, keep = int(keep64);
    DeviceGuard guard_(device);
    NCU(guard_.err);
    std::vector<double> yt(size_t(n) * m);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < m; ++k) yt[size_t(k) * n + i] = h_F[size_t(i) * m + k];
    NBuf bYT, bRank, bCnt, bLast, bCrowd, bSel;
    NCU(cudaMalloc(&bYT.p, yt.size() * sizeof(double)));
    NCU(cudaMalloc(&bRank.p, size_t(n) * sizeof(int)));
    NCU(cudaMalloc(&bCnt.p, rank_scratch_ints(n) * sizeof(int)));
    NCU(cudaMalloc(&bLast.p, sizeof(int)));
    NCU(cudaMalloc(&bCrowd.p, size_t(n) * sizeof(double)));
    NCU(cudaMalloc(&bSel.p, size_t(keep) * sizeof(int)));
    NCU(cudaMemcpy(bYT.p, yt.data(), yt.size() * sizeof(double), cudaMemcpyHostToDevice));
    NCHK(launch_rank(m, bYT.as<double>(), n, bRank.as<int>(), bCnt.as<int>(), keep, bLast.as<int>(), nullptr));
    NCHK(set_survival_smem(n));
    SurvivalScratch scratch;
    NCHK(scratch.alloc(n, m));
    SelectArgs sa{};
    sa.tab = scratch.tables(bYT.as<double>(), bRank.as<int>(), bLast.as<int>(), n, m);
    crowd_pos_kernel<<<(n + 31) / 32, 32 * CW_WARPS, crowd_smem(n)>>>(sa.tab);
    sa.crowd = bCrowd.as<double>();
    sa.n_cur = n; sa.keep = keep; sa.d = 0; sa.sel = bSel.as<int>();
    select_kernel<<<(n + 31) / 32, 32 * CW_WARPS, crowd_smem(n)>>>(sa);
    NCU(cudaGetLastError());
    NCU(cudaMemcpy(h_rank, bRank.p, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost));
    NCU(cudaMemcpy(h_crowd, bCrowd.p, size_t(n) * sizeof(double), cudaMemcpyDeviceToHost));
    NCU(cudaMemcpy(h_sel, bSel.p, size_t(keep) * sizeof(int), cudaMemcpyDeviceToHost));
    return AOED_OK;
}

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
