// Batched blocked Cholesky factorisation, triangular inverse and K^-1 on the FP64 tensor pipe (sm_100a).
//
// Replaces, for all (objective x restart) problems of one L-BFGS-B round at once, the O(N^3) part of sklearn's
// GaussianProcessRegressor.log_marginal_likelihood (`cholesky(K)`, `cho_solve((L, True), np.eye(N))`, _gpr.py:589-633,
// reached from autooed/mobo/surrogate_model/gp.py:139) and of the tail of `fit` (_gpr.py:349-367):
//
//   Cholesky  left-looking on 128-wide block columns.  Step k:  A[i][k] -= sum_{j<k} L[i][j] L[k][j]^T  for i >= k
//             (tgemm_kernel<NT>, k-depth 128 k), then the diagonal tile is factored AND inverted by one CTA per problem
//             (potrf_diag_kernel), then the panel is a GEMM:  L[i][k] = A[i][k] (L_kk^-1)^T  (tgemm_kernel<NT>, in place).
//   Inverse   recursive doubling on the tile grid: with X11, X22 the inverses of two adjacent diagonal blocks of 2^l tiles,
//             X21 = -X22 (L21 X11)  — two tile-GEMM launches per level, log2(T) levels (instead of T substitution steps).
//   K^-1      lower tiles of (L^-1)^T L^-1 — one launch.
//
// Every launch covers all problems of the batch: ticket -> (tile, problem), heaviest tiles first.
#include "blocked.h"

#include <algorithm>
#include <string>

#include "potrf_diag.cuh"
#include "tgemm.cuh"

namespace aoed {
namespace {

int bfail(const std::string& msg) { return aoed_detail_fail(-2 /*AOED_ERR_CUDA*/, msg.c_str()); }

#define BCU(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return bfail(std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" +    \
                         std::to_string(__LINE__));                                                       \
    } while (0)

template <bool NT>
int launch_tgemm(const CUtensorMap& mapA, const CUtensorMap& mapB, const BlockedPlan& plan, const BlockedSeg& seg, int batch,
                 int Npad, double* Out, int* counter, int sms, cudaStream_t s, int64_t* n_launches) {
    if (seg.n == 0) return 0;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        BCU(cudaFuncSetAttribute(tgemm_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PT_SMEM)));
        attr_set[dev & 63] = true;
    }
    TgemmArgs a;
    a.tiles = plan.dTiles + seg.off;
    a.ntiles = seg.n;
    a.batch = batch;
    a.rowsPerProblem = Npad;
    a.Out = Out;
    a.strideOut = int64_t(Npad) * Npad;
    a.ldo = Npad;
    a.counter = counter;
    a.dbg = nullptr;
    const int grid = std::min(seg.n * batch, sms);
    tgemm_kernel<NT><<<grid, PT_THREADS, PT_SMEM, s>>>(mapA, mapB, a);
    BCU(cudaGetLastError());
    if (n_launches) ++*n_launches;
    return 0;
}

}  // namespace

int blocked_plan_build(BlockedPlan& plan, int T) {
    blocked_plan_free(plan);
    plan.T = T;
    std::vector<TileDesc> t;
    const int KB = TM / TK;  // k-steps per tile
    auto sorted_seg = [&](size_t begin) {
        std::stable_sort(t.begin() + begin, t.end(), [](const TileDesc& x, const TileDesc& y) { return x.KT > y.KT; });
        BlockedSeg s;
        s.off = int(begin);
        s.n = int(t.size() - begin);
        return s;
    };
    plan.upd.assign(T, BlockedSeg());
    plan.pan.assign(T, BlockedSeg());
    for (int k = 0; k < T; ++k) {
        if (k > 0) {
            const size_t b0 = t.size();
            for (int i = k; i < T; ++i) t.push_back(TileDesc{i * TM, 0, k * TM, 0, k * KB, TG_SUB, i * TM, k * TM});
            plan.upd[k] = sorted_seg(b0);
        }
        if (k + 1 < T) {
            const size_t b0 = t.size();
            for (int i = k + 1; i < T; ++i) t.push_back(TileDesc{i * TM, k * TM, k * TM, k * TM, KB, TG_WRITE, i * TM, k * TM});
            plan.pan[k] = sorted_seg(b0);
        }
    }
    for (int s = 1; s < T; s *= 2) {
        const size_t a0 = t.size();
        for (int c0 = 0; c0 < T; c0 += 2 * s) {
            const int c1 = std::min(c0 + s, T), r0 = c1, r1 = std::min(r0 + s, T);
            for (int r = r0; r < r1; ++r)
                for (int c = c0; c < c1; ++c)  // P[r][c] = sum_{k = c .. c1-1} L[r][k] X[k][c]
                    t.push_back(TileDesc{r * TM, c * TM, c * TM, c * TM, (c1 - c) * KB, TG_WRITE, r * TM, c * TM});
        }
        plan.invA.push_back(sorted_seg(a0));
        const size_t b0 = t.size();
        for (int c0 = 0; c0 < T; c0 += 2 * s) {
            const int c1 = std::min(c0 + s, T), r0 = c1, r1 = std::min(r0 + s, T);
            for (int r = r0; r < r1; ++r)
                for (int c = c0; c < c1; ++c)  // X[r][c] = -sum_{k = r0 .. r} X[r][k] P[k][c]
                    t.push_back(TileDesc{r * TM, r0 * TM, r0 * TM, c * TM, (r - r0 + 1) * KB, TG_NEG, r * TM, c * TM});
        }
        plan.invB.push_back(sorted_seg(b0));
    }
    {
        const size_t b0 = t.size();
        for (int mt = 0; mt < T; ++mt)
            for (int ct = 0; ct <= mt; ++ct)  // Kinv[mt][ct] = sum_{k >= mt} XT[mt][k] X[k][ct]
                t.push_back(TileDesc{mt * TM, mt * TM, mt * TM, ct * TM, (T - mt) * KB, TG_WRITE, mt * TM, ct * TM});
        plan.lauum = sorted_seg(b0);
    }
    plan.n_counters = 2 * T + 2 * int(plan.invA.size()) + 1;
    BCU(cudaMalloc(&plan.dTiles, std::max<size_t>(t.size(), 1) * sizeof(TileDesc)));
    BCU(cudaMemcpy(plan.dTiles, t.data(), t.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
    return 0;
}

void blocked_plan_free(BlockedPlan& plan) {
    if (plan.dTiles) cudaFree(plan.dTiles);
    plan = BlockedPlan();
}

int blocked_mats_alloc(BlockedMats& m, int Npad, int cap) {
    blocked_mats_free(m);
    const size_t bytes = size_t(cap) * Npad * Npad * sizeof(double);
    BCU(cudaMalloc(&m.K, bytes));
    BCU(cudaMalloc(&m.X, bytes));
    BCU(cudaMalloc(&m.XT, bytes));
    BCU(cudaMalloc(&m.P, bytes));
    m.Npad = Npad;
    m.cap = cap;
    const int64_t rows = int64_t(cap) * Npad;
    int rc;
    if ((rc = make_tensor_map_2d(&m.K_a, m.K, rows, Npad, Npad, TM)) != 0) return rc;
    if ((rc = make_tensor_map_2d(&m.X_a, m.X, rows, Npad, Npad, TM)) != 0) return rc;
    if ((rc = make_tensor_map_2d(&m.X_b, m.X, rows, Npad, Npad, TK)) != 0) return rc;
    if ((rc = make_tensor_map_2d(&m.XT_a, m.XT, rows, Npad, Npad, TM)) != 0) return rc;
    if ((rc = make_tensor_map_2d(&m.P_b, m.P, rows, Npad, Npad, TK)) != 0) return rc;
    return 0;
}

void blocked_mats_free(BlockedMats& m) {
    double** ptrs[] = {&m.K, &m.X, &m.XT, &m.P};
    for (auto pp : ptrs) {
        if (*pp) cudaFree(*pp);
        *pp = nullptr;
    }
    m.Npad = m.cap = 0;
}

int blocked_chol_inverse(const BlockedPlan& plan, const BlockedMats& m, int batch, int* d_info, int* d_counters, int sms,
                         cudaStream_t s, int64_t* n_launches) {
    const int T = plan.T, Npad = m.Npad;
    if (T * TM != Npad || batch > m.cap) return aoed_detail_fail(-1, "blocked_chol_inverse: plan / buffers do not match");
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        BCU(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PD_SMEM)));
        attr_set[dev & 63] = true;
    }
    BCU(cudaMemsetAsync(d_counters, 0, size_t(plan.n_counters) * sizeof(int), s));
    int* ctr = d_counters;
    int rc;
    for (int k = 0; k < T; ++k) {
        if ((rc = launch_tgemm<true>(m.K_a, m.K_a, plan, plan.upd[k], batch, Npad, m.K, ctr++, sms, s, n_launches)) != 0) return rc;
        PotrfDiagArgs pa;
        pa.K = m.K;
        pa.Xinv = m.X;
        pa.stride = int64_t(Npad) * Npad;
        pa.ld = Npad;
        pa.k = k;
        pa.info = d_info;
        potrf_diag_kernel<<<batch, PD_THREADS, PD_SMEM, s>>>(pa);
        BCU(cudaGetLastError());
        if (n_launches) ++*n_launches;
        if ((rc = launch_tgemm<true>(m.K_a, m.X_a, plan, plan.pan[k], batch, Npad, m.K, ctr++, sms, s, n_launches)) != 0) return rc;
    }
    for (size_t l = 0; l < plan.invA.size(); ++l) {
        if ((rc = launch_tgemm<false>(m.K_a, m.X_b, plan, plan.invA[l], batch, Npad, m.P, ctr++, sms, s, n_launches)) != 0) return rc;
        if ((rc = launch_tgemm<false>(m.X_a, m.P_b, plan, plan.invB[l], batch, Npad, m.X, ctr++, sms, s, n_launches)) != 0) return rc;
    }
    return 0;
}

int blocked_kinv(const BlockedPlan& plan, const BlockedMats& m, int batch, int* d_counters, int sms, cudaStream_t s,
                 int64_t* n_launches) {
    // the last counter of the sequence belongs to this launch
    int* ctr = d_counters + plan.n_counters - 1;
    BCU(cudaMemsetAsync(ctr, 0, sizeof(int), s));
    return launch_tgemm<false>(m.XT_a, m.X_b, plan, plan.lauum, batch, m.Npad, m.P, ctr, sms, s, n_launches);
}

}  // namespace aoed
