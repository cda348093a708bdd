// Host side of the C ABI for the GP surrogate: handle, device residency, launch orchestration.
// See include/aoed_b200.h for the contract of every exported function.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/aoed_b200.h"
#include "blocked.h"
#include "launchers.h"
#include "fit.cuh"
#include "fit_small.cuh"
#include "fit_small2.cuh"
#include "fit_launch.h"
#include "hessian.cuh"
#include "posterior.cuh"
#include "small_eval.cuh"
#include "trmm.cuh"
#include "trmm_tma.cuh"

using namespace aoed;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(AOED_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + \
                                           ":" + std::to_string(__LINE__));                                   \
    } while (0)

#define CHK(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_ != AOED_OK) return rc_; \
    } while (0)

int pick_dp(int d) {
    if (d <= 8) return 8;
    if (d <= 16) return 16;
    if (d <= 24) return 24;
    if (d <= 32) return 32;
    if (d <= 48) return 48;
    if (d <= 64) return 64;
    if (d <= 96) return 96;
    if (d <= 128) return 128;
    return -1;
}

struct Workspace {
    int64_t cols = 0;  // capacity in candidates (multiple of 128)
    double *XT = nullptr, *KsT = nullptr, *Gr = nullptr, *W = nullptr, *V = nullptr;
    double *sumsq = nullptr, *part = nullptr, *part2 = nullptr;
    // staging for the host-pointer API
    double *Xc = nullptr, *oF = nullptr, *oS = nullptr, *oA = nullptr, *odF = nullptr, *odS = nullptr, *odA = nullptr;
    double* hX = nullptr;  // pinned input staging
    // Hessian path (lazily allocated): |L^-1 dk_i|^2 partials, Gram matrices, internal first-order outputs, staging
    CUtensorMap mapKsT, mapW;  // TMA views of the two slabs that serve as B operand of the triangular products
    int64_t hrows = 0;
    double *sumsq2 = nullptr, *gram = nullptr, *bF = nullptr, *bS = nullptr, *bdF = nullptr, *bdS = nullptr;
    double *ohF = nullptr, *ohS = nullptr, *ohA = nullptr;
};

}  // namespace

struct aoed_gp {
    int device = 0, d = 0, m = 0, nu = 1, DP = 8, DE = 8;
    int N = 0, Npad = 0, Kpad = 0, mTiles = 0, JS = 1, jchunk = 0;  // JS: CAPACITY of the training-index split (partials)
    int cap_Npad = 0;  // row capacity the device buffers were allocated for (re-used while N stays in the same 128-bucket)
    int occ_xcov = 2, occ_gradsum = 5;  // resident CTAs per SM of the two thread-per-candidate kernels
    int blocked_cap_max = 1 << 30;       // problems per group the scratch budget allowed when it was last sized
    bool has_train = false, has_theta = false;
    std::vector<double> theta, y_mean, y_scale, lml, Xn, Yn;
    // device model state
    double *dX = nullptr, *dY = nullptr;  // [N][d], [m][Npad]
    double* dXT = nullptr;                // [d][Npad] dimension-major copy of dX, zero padded
    double *dTs = nullptr, *dAlpha = nullptr, *dEll = nullptr, *dLinv = nullptr, *dLinvT = nullptr, *dL = nullptr;
    ModelDev* dModel = nullptr;
    double* dC1C2 = nullptr;        // [m][2]
    CUtensorMap mapLinv, mapLinvT;  // TMA views of the cached factors (A operands)
    int* dTickets = nullptr;        // tile ticket counters of the persistent triangular products
    int* hDbg = nullptr;            // host-mapped diagnostics of the persistent kernel (AOED_TRMM_DEBUG=1)
    int ticket_slot = 0;
    int trmm_variant = 0;           // 0: persistent TMA kernel, 64 / 128: cp.async tiled kernel with that column tile
    int ssf = PT_SS_PER_TILE;       // sum-of-squares partials per 128-row tile (4 for the TMA kernel, 1 otherwise)
    // blocked factorisation (blocked.cu): tile plan of the current matrix size, one batch of matrices, per-problem vectors
    BlockedPlan plan;
    BlockedMats mats;
    double *bTs = nullptr, *bC1C2 = nullptr, *bZ = nullptr, *bAlpha = nullptr, *bPartial = nullptr;
    int *bInfo = nullptr, *bCounters = nullptr;
    double *dThetaB = nullptr, *dResB = nullptr;  // batched LML inputs / results
    int* dObjB = nullptr;
    int batchCap = 0;
    int64_t n_fit_small = 0, n_fit_blocked = 0;
    Workspace ws[2];
    cudaStream_t streams[2] = {nullptr, nullptr};  // [0]: compute (all kernels, in order), [1]: host -> device copies
    cudaStream_t d2h_stream = nullptr;             // device -> host copies
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // counters
    int64_t n_launches = 0, trmm_launches = 0;
    double trmm_ms = 0.0, trmm_flops = 0.0;
    bool profiling = false;
};

namespace {

void free_ws(Workspace& w) {
    double** ptrs[] = {&w.XT, &w.KsT, &w.Gr, &w.W, &w.V, &w.sumsq, &w.part, &w.part2,
                       &w.Xc, &w.oF, &w.oS, &w.oA, &w.odF, &w.odS, &w.odA,
                       &w.sumsq2, &w.gram, &w.bF, &w.bS, &w.bdF, &w.bdS, &w.ohF, &w.ohS, &w.ohA};
    for (auto pp : ptrs) {
        if (*pp) cudaFree(*pp);
        *pp = nullptr;
    }
    if (w.hX) cudaFreeHost(w.hX);
    w.hX = nullptr;
    w.cols = 0;
    w.hrows = 0;
}

void free_blocked(aoed_gp* gp) {
    blocked_mats_free(gp->mats);
    blocked_plan_free(gp->plan);
    double** ptrs[] = {&gp->bTs, &gp->bC1C2, &gp->bZ, &gp->bAlpha, &gp->bPartial};
    for (auto pp : ptrs) {
        if (*pp) cudaFree(*pp);
        *pp = nullptr;
    }
    if (gp->bInfo) cudaFree(gp->bInfo);
    if (gp->bCounters) cudaFree(gp->bCounters);
    gp->bInfo = gp->bCounters = nullptr;
}

void free_model(aoed_gp* gp) {
    double** ptrs[] = {&gp->dX, &gp->dXT, &gp->dY, &gp->dTs, &gp->dAlpha, &gp->dEll, &gp->dLinv, &gp->dLinvT, &gp->dL, &gp->dC1C2};
    for (auto pp : ptrs) {
        if (*pp) cudaFree(*pp);
        *pp = nullptr;
    }
    if (gp->dModel) cudaFree(gp->dModel);
    gp->dModel = nullptr;
    free_blocked(gp);
    gp->cap_Npad = 0;
    if (gp->dThetaB) cudaFree(gp->dThetaB);
    if (gp->dResB) cudaFree(gp->dResB);
    if (gp->dObjB) cudaFree(gp->dObjB);
    if (gp->dTickets) cudaFree(gp->dTickets);
    gp->dTickets = nullptr;
    gp->dThetaB = gp->dResB = nullptr;
    gp->dObjB = nullptr;
    gp->batchCap = 0;
    free_ws(gp->ws[0]);
    free_ws(gp->ws[1]);
}

// Candidates per chunk (slab columns).  Larger chunks amortise launch boundaries of the persistent kernels (measured:
// 94.5 % / 95.6 % / 96.2 % of DGEMM peak at 4096 / 8192 / 16384); the host-pointer API prefers finer chunks because H2D,
// compute and D2H of consecutive chunks overlap.  AOED_CHUNK overrides both.
int64_t chunk_cap(bool device_resident = false) {
    const char* e = getenv("AOED_CHUNK");
    int64_t v = e ? atoll(e) : (device_resident ? 16384 : 8192);
    if (v < 128) v = 128;
    return round_up64(v, 128);
}

int make_map_2d(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

int ensure_ws(aoed_gp* gp, Workspace& w, int64_t cols, bool staging) {
    cols = round_up64(cols, 128);
    const int m = gp->m, DP = gp->DP, Npad = gp->Npad;
    if (w.cols < cols) {
        free_ws(w);
        const size_t slab = size_t(m) * Npad * cols * sizeof(double);
        CU(cudaMalloc(&w.XT, size_t(DP) * cols * sizeof(double)));
        CU(cudaMalloc(&w.KsT, slab));
        CU(cudaMalloc(&w.Gr, slab));
        CU(cudaMalloc(&w.W, slab));
        CU(cudaMalloc(&w.V, slab));
        CU(cudaMemset(w.KsT, 0, slab));
        CU(cudaMemset(w.Gr, 0, slab));
        CU(cudaMemset(w.W, 0, slab));
        CU(cudaMemset(w.V, 0, slab));
        CU(cudaMalloc(&w.sumsq, size_t(m) * gp->mTiles * gp->ssf * cols * sizeof(double)));
        CU(cudaMalloc(&w.part, size_t(m) * gp->JS * (2 + DP) * cols * sizeof(double)));
        // gradient moments of the variance: one slice per training-index split (gradsum_kernel) or per 32-row warp slice of the
        // UPPER product when they are fused into its epilogue
        CU(cudaMalloc(&w.part2, size_t(m) * std::max(gp->JS, gp->mTiles * gp->ssf) * (1 + DP) * cols * sizeof(double)));
        if (gp->trmm_variant == 0) {
            CHK(make_map_2d(&w.mapKsT, w.KsT, int64_t(m) * Npad, cols, cols, TK));
            CHK(make_map_2d(&w.mapW, w.W, int64_t(m) * Npad, cols, cols, TK));
        }
        w.cols = cols;
    }
    if (staging && w.Xc == nullptr) {
        const size_t c = size_t(w.cols);
        CU(cudaMalloc(&w.Xc, c * gp->d * sizeof(double)));
        CU(cudaMalloc(&w.oF, c * m * sizeof(double)));
        CU(cudaMalloc(&w.oS, c * m * sizeof(double)));
        CU(cudaMalloc(&w.oA, c * m * sizeof(double)));
        CU(cudaMalloc(&w.odF, c * m * gp->d * sizeof(double)));
        CU(cudaMalloc(&w.odS, c * m * gp->d * sizeof(double)));
        CU(cudaMalloc(&w.odA, c * m * gp->d * sizeof(double)));
        CU(cudaMallocHost(&w.hX, c * gp->d * sizeof(double)));
    }
    return AOED_OK;
}

// Training-index split for the thread-per-candidate kernels: the grid is colTiles x JS x m CTAs; pick JS (<= capacity) so
// that the CTA count fills whole waves of (SMs x resident CTAs) — a 2.6-wave grid runs as 3 waves, a 1.04-wave grid as 2.
void choose_split(int N, int colTiles, int m, int js_cap, int ctas_per_wave, int& JS, int& jchunk) {
    double best = -1.0;
    JS = 1;
    const int js_max = std::max(1, std::min(js_cap, (N + XC_JT - 1) / XC_JT));
    for (int js = 1; js <= js_max; ++js) {
        const int jc = round_up((N + js - 1) / js, XC_JT);
        const int real = (N + jc - 1) / jc;
        if (real != js) continue;
        const double ctas = double(colTiles) * js * m;
        const double waves = std::ceil(ctas / ctas_per_wave);
        double eff = ctas / (waves * ctas_per_wave);  // fraction of the last wave's slots that do work
        if (waves > 1.0) eff = std::min(eff + 0.02 * (waves - 1.0), 1.0);  // more waves amortise a ragged tail
        if (eff > best + 1e-9) {
            best = eff;
            JS = js;
        }
    }
    jchunk = round_up((N + JS - 1) / JS, XC_JT);
}

template <bool UPPER, int TNV>
int launch_trmm_cfg(const TrmmArgs& a, int colTiles128, int batch, cudaStream_t s) {
    static bool attr_set[64] = {false};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        CU(cudaFuncSetAttribute(trmm_kernel<UPPER, TNV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                int(TrmmCfg<TNV>::SMEM)));
        attr_set[dev & 63] = true;
    }
    TrmmArgs args = a;
    args.colTiles = colTiles128 * (TN / TNV);
    dim3 grid(args.colTiles * batch, a.mTiles, 1);
    trmm_kernel<UPPER, TNV><<<grid, TrmmCfg<TNV>::THREADS, TrmmCfg<TNV>::SMEM, s>>>(args);
    return AOED_OK;
}

// ---- TMA descriptors ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// row-major FP64 matrix [rows][ld] (cols valid) viewed through boxes of box_rows x 16 columns, 128-byte swizzle
int make_map_2d(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(AOED_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * sizeof(double)};
    cuuint32_t box[2] = {16u, cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AOED_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
    return AOED_OK;
}

}  // namespace
namespace aoed {
int make_tensor_map_2d(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    return make_map_2d(map, base, rows, cols, ld, box_rows);
}
}  // namespace aoed
namespace {

int trmm_variant_from_env() {
    const char* e = getenv("AOED_TRMM");
    if (e && strcmp(e, "tile128") == 0) return 128;
    if (e && strcmp(e, "tile64") == 0) return 64;
    return 0;
}

int sm_count(int device) {
    static int n[64] = {0};
    if (device < 0 || device >= 64) return 148;
    if (n[device] == 0) {
        cudaDeviceProp prop;
        n[device] = (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ? prop.multiProcessorCount : 148;
    }
    return n[device];
}

// `ticket`: the launch's tile counter; it must not be shared with a launch that can still be running on ANOTHER stream
// (launches on one stream are ordered, so a per-stream counter or ring is safe).
// optional fused epilogue of the UPPER product (variance-gradient moments, see TrmmTmaArgs)
struct FusedMoments {
    const double* Gr = nullptr;
    const double* Ts = nullptr;
    double* part2 = nullptr;
    int DPs = 0, dims = 0;
};

template <bool UPPER>
int launch_trmm_tma(aoed_gp* gp, const CUtensorMap& mapA, const CUtensorMap& mapB, const TrmmArgs& a, int colTiles,
                    int batch, cudaStream_t s, int* ticket = nullptr, const FusedMoments* fm = nullptr) {
    static bool attr_set[64] = {false};  // function attributes are per device
    const int dev = gp->device & 63;
    if (!attr_set[dev]) {
        CU(cudaFuncSetAttribute(trmm_tma_kernel<UPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(PT_SMEM)));
        attr_set[dev] = true;
    }
    TrmmTmaArgs t;
    t.Out = a.Out; t.strideOut = a.strideOut; t.ldo = a.ldo;
    t.sumsq = a.sumsq; t.strideSS = a.strideSS;
    t.M = a.M; t.Kpad = a.Kpad; t.Npad = gp->Npad; t.mTiles = a.mTiles; t.colTiles = colTiles; t.batch = batch;
    t.Gr = fm ? fm->Gr : nullptr; t.Ts = fm ? fm->Ts : nullptr; t.part2 = fm ? fm->part2 : nullptr;
    t.DPs = fm ? fm->DPs : 0; t.dims = fm ? fm->dims : 0;
    if (ticket == nullptr) {  // evaluation path: ring on the (single) compute stream
        gp->ticket_slot = (gp->ticket_slot + 1) % 64;
        ticket = gp->dTickets + gp->ticket_slot;
    }
    t.counter = ticket;
    static const int col_group = [] {
        const char* e = getenv("AOED_TRMM_COLGROUP");
        const int v = e ? atoi(e) : PT_COL_GROUP;
        return v >= 1 ? v : PT_COL_GROUP;
    }();
    t.colGroup = col_group;
    CU(cudaMemsetAsync(t.counter, 0, sizeof(int), s));
    static const bool debug = getenv("AOED_TRMM_DEBUG") != nullptr;
    if (debug && gp->hDbg == nullptr) {
        CU(cudaHostAlloc(&gp->hDbg, 64 * sizeof(int), cudaHostAllocMapped));
        memset(gp->hDbg, 0, 64 * sizeof(int));
    }
    t.dbg = debug ? gp->hDbg : nullptr;
    const int tiles = a.mTiles * colTiles * batch;
    const int grid = std::min(tiles, sm_count(gp->device));
    trmm_tma_kernel<UPPER><<<grid, PT_THREADS, PT_SMEM, s>>>(mapA, mapB, t);
    if (debug) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            const int* d = gp->hDbg;
            return fail(AOED_ERR_CUDA, std::string("trmm_tma_kernel<") + (UPPER ? "U" : "L") + "> failed: " + cudaGetErrorString(e) +
                                           " dbg where=" + std::to_string(d[0]) + " block=" + std::to_string(d[1]) + " thread=" +
                                           std::to_string(d[2]) + " it=" + std::to_string(d[3]) + " parity=" + std::to_string(d[4]) +
                                           " grid=" + std::to_string(grid) + " tiles=" + std::to_string(tiles));
        }
    }
    return AOED_OK;
}

// `colTiles` counts 128-column groups of the slab; mapB must view the slab passed as a.B
int launch_trmm(aoed_gp* gp, bool upper, const TrmmArgs& a, const CUtensorMap& mapB, int colTiles, int batch,
                cudaStream_t s, const FusedMoments* fm = nullptr) {
    if (gp->profiling) CU(cudaEventRecord(gp->ev0, s));
    if (gp->trmm_variant == 0) {
        if (upper) CHK(launch_trmm_tma<true>(gp, gp->mapLinvT, mapB, a, colTiles, batch, s, nullptr, fm));
        else CHK(launch_trmm_tma<false>(gp, gp->mapLinv, mapB, a, colTiles, batch, s));
    } else if (gp->trmm_variant == 128) {
        if (upper) CHK((launch_trmm_cfg<true, 128>(a, colTiles, batch, s)));
        else CHK((launch_trmm_cfg<false, 128>(a, colTiles, batch, s)));
    } else {
        if (upper) CHK((launch_trmm_cfg<true, 64>(a, colTiles, batch, s)));
        else CHK((launch_trmm_cfg<false, 64>(a, colTiles, batch, s)));
    }
    CU(cudaGetLastError());
    gp->n_launches++;
    gp->trmm_launches++;
    // useful flops: one triangular product of an N x N factor with (colTiles*128) columns
    gp->trmm_flops += double(batch) * double(gp->N) * (double(gp->N) + 1.0) * double(colTiles) * TN;
    if (gp->profiling) {
        CU(cudaEventRecord(gp->ev1, s));
        CU(cudaEventSynchronize(gp->ev1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, gp->ev0, gp->ev1));
        gp->trmm_ms += ms;
    }
    return AOED_OK;
}

struct EvalOut {
    double *F, *S, *dF, *dS, *A, *dA;
};

// value-only calls on small batches and small training sets: one fused kernel (small_eval.cuh).
// AOED_SMALL_EVAL=0 sends them down the chunked path instead (tests compare the two).
bool small_eval_applies(const aoed_gp* gp, int64_t rows, bool want_grad) {
    if (want_grad || gp->Npad > SE_MAX_NPAD || rows > SE_MAX_ROWS) return false;
    if (small_eval_smem(gp->Npad, gp->DP) > 227 * 1024) return false;  // wide designs on 256 training rows: the chunked kernels
    const char* env = getenv("AOED_SMALL_EVAL");
    return !(env && env[0] == '0');
}

int eval_small(aoed_gp* gp, const double* d_X, int64_t rows, bool want_std, const AcqDev& acq, const EvalOut& out, cudaStream_t s) {
    SmallEvalArgs a;
    a.X = d_X; a.Ts = gp->dTs; a.alpha = gp->dAlpha; a.ell = gp->dEll; a.model = gp->dModel; a.LinvT = gp->dLinvT;
    a.N = gp->N; a.Npad = gp->Npad; a.d = gp->d; a.m = gp->m; a.std = want_std ? 1 : 0; a.rows = rows; a.acq = acq;
    a.F = out.F; a.S = out.S; a.A = out.A;
    const dim3 grid(unsigned((rows + SE_TC - 1) / SE_TC), gp->m);
    const size_t smem = small_eval_smem(gp->Npad, gp->DP);
    launch_small_eval_any(gp->nu, gp->DP, a, grid, smem, s);
    gp->n_launches++;
    CU(cudaGetLastError());
    return AOED_OK;
}

// One chunk of `rows` candidates (rows <= ws.cols), device-resident, row-major input d_X[rows][d].
int eval_chunk(aoed_gp* gp, Workspace& w, const double* d_X, int64_t rows, bool want_std, bool want_grad,
               const AcqDev& acq, const EvalOut& out, cudaStream_t s, bool keep_V = false) {
    if (small_eval_applies(gp, rows, want_grad)) return eval_small(gp, d_X, rows, want_std, acq, out, s);
    const int m = gp->m, DP = gp->DP, Npad = gp->Npad;
    const int ldc = int(w.cols);
    const int colTiles = int((rows + 127) / 128);
    const int64_t slab = int64_t(Npad) * ldc;
    {
        const int cover = colTiles * 128;
        transpose_x_kernel<<<(cover + 127) / 128, 128, 0, s>>>(d_X, rows, gp->d, DP, ldc, w.XT);
        gp->n_launches++;
    }
    XcovArgs xa;
    xa.XT = w.XT; xa.Ts = gp->dTs; xa.alpha = gp->dAlpha; xa.ell = gp->dEll; xa.model = gp->dModel;
    xa.KsT = want_std ? w.KsT : nullptr;
    xa.Gr = (want_std && want_grad) ? w.Gr : nullptr;
    xa.part = w.part; xa.strideSlab = slab;
    const int sms = sm_count(gp->device);
    int JS1, jchunk1, JS2 = 1, jchunk2 = gp->jchunk;
    choose_split(gp->N, colTiles, m, gp->JS, sms * gp->occ_xcov, JS1, jchunk1);
    xa.N = gp->N; xa.Npad = Npad; xa.ldc = ldc; xa.JS = JS1; xa.jchunk = jchunk1;
    launch_xcov_any(gp->nu, DP, gp->DE, xa, dim3(colTiles, JS1, m), want_grad, s);
    gp->n_launches++;
    CU(cudaGetLastError());
    if (want_std) {
        TrmmArgs t;
        t.A = gp->dLinv; t.strideA = int64_t(Npad) * Npad; t.lda = Npad;
        t.B = w.KsT; t.strideB = slab; t.ldb = ldc;
        t.Out = want_grad ? w.W : nullptr; t.strideOut = slab; t.ldo = ldc;
        t.sumsq = w.sumsq; t.strideSS = int64_t(gp->mTiles) * gp->ssf * ldc;
        t.M = gp->N; t.Kpad = gp->Kpad; t.mTiles = gp->mTiles;
        CHK(launch_trmm(gp, false, t, w.mapKsT, colTiles, m, s));
        if (want_grad) {
            TrmmArgs u = t;
            u.A = gp->dLinvT; u.B = w.W; u.Out = w.V; u.sumsq = nullptr;
            // AOED_GRADSUM=fused forms the moments sum_j Gr_j v_j [ts_jk] from the accumulators in the epilogue of the persistent
            // UPPER product (v never leaves the SM, no gradsum launch, 1.4 GB less HBM traffic per 16384-candidate chunk).
            // Parity-tested, but NOT the default: measured at c4 it costs +11 % on the UPPER launches (6.15 instead of 5.83 ms
            // average launch, step 852 instead of 834 ms) — all 16 consumer warps reach their epilogues together and the
            // 21 dependent load -> FMA -> shuffle rounds of the slowest warp stall the whole TMA ring, which is worth more
            // than the 28 ms the separate HBM-bound kernel takes (profiles/bench_fused_gradsum_r02.json).
            static const bool fused = getenv("AOED_GRADSUM") && strcmp(getenv("AOED_GRADSUM"), "fused") == 0;
            if (gp->trmm_variant == 0 && fused) {
                FusedMoments fm;
                fm.Gr = w.Gr; fm.Ts = gp->dTs; fm.part2 = w.part2; fm.DPs = DP; fm.dims = gp->d;
                if (!keep_V) u.Out = nullptr;
                CHK(launch_trmm(gp, true, u, w.mapW, colTiles, m, s, &fm));
                JS2 = gp->mTiles * gp->ssf;
            } else {
                CHK(launch_trmm(gp, true, u, w.mapW, colTiles, m, s));
                GradsumArgs ga;
                ga.V = w.V; ga.Gr = w.Gr; ga.Ts = gp->dTs; ga.part2 = w.part2; ga.strideSlab = slab;
                choose_split(gp->N, colTiles, m, gp->JS, sms * gp->occ_gradsum, JS2, jchunk2);
                ga.N = gp->N; ga.Npad = Npad; ga.ldc = ldc; ga.JS = JS2; ga.jchunk = jchunk2;
                launch_gradsum(DP, gp->DE, ga, dim3(colTiles, JS2, m), s);
                gp->n_launches++;
                CU(cudaGetLastError());
            }
        }
    }
    EpilogueArgs e;
    e.XT = w.XT; e.ell = gp->dEll; e.model = gp->dModel; e.part = w.part; e.part2 = w.part2; e.sumsq = w.sumsq;
    e.F = out.F; e.S = out.S; e.dF = out.dF; e.dS = out.dS; e.A = out.A; e.dA = out.dA;
    e.rows = rows; e.m = m; e.d = gp->d; e.DP = DP; e.ldc = ldc; e.JS = JS1; e.JS2 = JS2; e.mTiles = gp->mTiles * gp->ssf;
    e.std = want_std; e.grad = want_grad; e.acq = acq;
    epilogue_kernel<<<dim3(unsigned((rows + EPI_CX - 1) / EPI_CX), m), dim3(EPI_CX, EPI_KY), 0, s>>>(e);
    gp->n_launches++;
    CU(cudaGetLastError());
    return AOED_OK;
}

int make_acq(const aoed_gp* gp, int kind, uint32_t flags, const double* param, AcqDev& acq) {
    memset(&acq, 0, sizeof(acq));
    acq.kind = kind;
    acq.exact = (flags & AOED_EVAL_EXACT) ? 1 : 0;
    if (kind < AOED_ACQ_NONE || kind > AOED_ACQ_PI) return fail(AOED_ERR_INVALID, "unknown acquisition kind");
    if (kind == AOED_ACQ_UCB) {
        if (!param || param[0] < 1.0) return fail(AOED_ERR_INVALID, "UCB needs n_sample >= 1");
        acq.lam = sqrt(log(param[0]) / param[0]);
    } else if (kind == AOED_ACQ_EI || kind == AOED_ACQ_PI) {
        if (!param) return fail(AOED_ERR_INVALID, "EI/PI need y_min");
        if (gp->m > 16) return fail(AOED_ERR_UNSUPPORTED, "EI/PI support n_obj <= 16");
        for (int o = 0; o < gp->m; ++o) acq.y_min[o] = param[o];
    }
    return AOED_OK;
}

// rows of candidates per chunk on the Hessian path: the d k*/dx_i right-hand sides take rows*d slab columns
int64_t hess_rows_cap(const aoed_gp* gp, int64_t n_rows) {
    const int64_t by_cols = std::max<int64_t>(1, chunk_cap() / gp->d);
    return std::max<int64_t>(1, std::min<int64_t>(n_rows, by_cols));
}

int ensure_ws_hess(aoed_gp* gp, Workspace& w, int64_t hrows, bool staging) {
    CHK(ensure_ws(gp, w, round_up64(hrows * gp->d, 128), staging));
    const size_t m = gp->m, d = gp->d;
    if (w.hrows < hrows) {
        double** ptrs[] = {&w.sumsq2, &w.gram, &w.bF, &w.bS, &w.bdF, &w.bdS, &w.ohF, &w.ohS, &w.ohA};
        for (auto pp : ptrs) {
            if (*pp) cudaFree(*pp);
            *pp = nullptr;
        }
        const size_t r = size_t(hrows);
        CU(cudaMalloc(&w.sumsq2, m * gp->mTiles * gp->ssf * size_t(w.cols) * sizeof(double)));
        CU(cudaMalloc(&w.gram, m * r * d * d * sizeof(double)));
        CU(cudaMalloc(&w.bF, r * m * sizeof(double)));
        CU(cudaMalloc(&w.bS, r * m * sizeof(double)));
        CU(cudaMalloc(&w.bdF, r * m * d * sizeof(double)));
        CU(cudaMalloc(&w.bdS, r * m * d * sizeof(double)));
        w.hrows = hrows;
    }
    if (staging && w.ohF == nullptr) {
        const size_t n = size_t(w.hrows) * m * d * d * sizeof(double);
        CU(cudaMalloc(&w.ohF, n));
        CU(cudaMalloc(&w.ohS, n));
        CU(cudaMalloc(&w.ohA, n));
    }
    return AOED_OK;
}

struct HessOut {
    double *hF, *hS, *hA;
};

// One chunk with second derivatives: first-order pipeline (always with gradients: the variance Hessian and the
// EI / PI Hessians need them, gp.py:236-241), then the d k*/dx_i triangular product and the fused Hessian kernel.
int eval_hess_chunk(aoed_gp* gp, Workspace& w, const double* d_X, int64_t rows, bool want_std, const AcqDev& acq,
                    const EvalOut& out, const HessOut& hout, cudaStream_t s) {
    const int m = gp->m, d = gp->d, DP = gp->DP, Npad = gp->Npad;
    const int ldc = int(w.cols);
    const int64_t slab = int64_t(Npad) * ldc;
    EvalOut in = out;  // first-order results are inputs of the Hessian kernel: route through internal buffers if unwanted
    if (!in.F) in.F = w.bF;
    if (!in.dF) in.dF = w.bdF;
    if (want_std) {
        if (!in.S) in.S = w.bS;
        if (!in.dS) in.dS = w.bdS;
    }
    CHK(eval_chunk(gp, w, d_X, rows, want_std, true, acq, in, s, /*keep_V=*/true));
    const bool exact = acq.exact != 0;
    if (want_std) {
        DkcolArgs da;
        da.XT = w.XT; da.Ts = gp->dTs; da.ell = gp->dEll; da.Gr = w.Gr; da.DK = w.KsT; da.strideSlab = slab;
        da.rows = int(rows); da.d = d; da.DP = DP; da.N = gp->N; da.Npad = Npad; da.ldc = ldc; da.ldx = ldc;
        da.JS = gp->JS; da.jchunk = gp->jchunk;
        const int colTiles = int((rows * d + 127) / 128);
        dkcol_kernel<<<dim3(colTiles, gp->JS, m), 128, 0, s>>>(da);
        gp->n_launches++;
        CU(cudaGetLastError());
        TrmmArgs t;
        t.A = gp->dLinv; t.strideA = int64_t(Npad) * Npad; t.lda = Npad;
        t.B = w.KsT; t.strideB = slab; t.ldb = ldc;
        t.Out = exact ? w.W : nullptr; t.strideOut = slab; t.ldo = ldc;
        t.sumsq = exact ? nullptr : w.sumsq2; t.strideSS = int64_t(gp->mTiles) * gp->ssf * ldc;
        t.M = gp->N; t.Kpad = gp->Kpad; t.mTiles = gp->mTiles;
        CHK(launch_trmm(gp, false, t, w.mapKsT, colTiles, m, s));
        if (exact) {
            GramArgs ga;
            ga.W = w.W; ga.G = w.gram; ga.strideSlab = slab; ga.rows = int(rows); ga.d = d; ga.N = gp->N; ga.ldc = ldc;
            gram_kernel<<<dim3(unsigned(rows), m), HS_THREADS, 0, s>>>(ga);
            gp->n_launches++;
            CU(cudaGetLastError());
        }
    }
    HessArgs h;
    h.XT = w.XT; h.Ts = gp->dTs; h.ell = gp->dEll; h.alpha = gp->dAlpha; h.model = gp->dModel;
    h.V = w.V; h.sumsq = w.sumsq; h.mid_ss = w.sumsq2; h.gram = w.gram;
    h.F = in.F; h.S = in.S; h.dF = in.dF; h.dS = in.dS;
    h.hF = hout.hF; h.hS = want_std ? hout.hS : nullptr; h.hA = hout.hA;
    h.strideSlab = slab; h.rows = int(rows); h.m = m; h.d = d; h.N = gp->N; h.Npad = Npad; h.ldc = ldc; h.ldx = ldc;
    h.mTiles = gp->mTiles * gp->ssf; h.std = want_std ? 1 : 0; h.acq = acq;
    launch_hess(gp->nu, DP, h, dim3(unsigned(rows), m), s);
    gp->n_launches++;
    CU(cudaGetLastError());
    return AOED_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int aoed_detail_fail(int code, const char* msg) { return fail(code, msg ? msg : ""); }

int aoed_abi_version(void) { return AOED_ABI_VERSION; }
const char* aoed_last_error(void) { return g_err.c_str(); }

int aoed_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int aoed_host_alloc(void** out, uint64_t bytes) {
    if (!out) return fail(AOED_ERR_INVALID, "null out");
    CU(cudaMallocHost(out, bytes ? bytes : 8));
    return AOED_OK;
}

int aoed_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return AOED_OK;
}

int aoed_gp_create(aoed_gp_t** out, int device, int n_var, int n_obj, int nu) {
    if (!out) return fail(AOED_ERR_INVALID, "null out");
    if (aoed_device_count() <= device || device < 0)
        return fail(AOED_ERR_NO_DEVICE, "no CUDA device " + std::to_string(device) + " (there is no CPU fallback)");
    if (n_var < 1 || n_obj < 1) return fail(AOED_ERR_INVALID, "n_var and n_obj must be positive");
    if (!(nu == 1 || nu == 3 || nu == 5 || nu == -1)) return fail(AOED_ERR_INVALID, "nu must be 1, 3, 5 or -1");
    const int DP = pick_dp(n_var);
    if (DP < 0) return fail(AOED_ERR_UNSUPPORTED, "n_var > 128 is not supported");
    DeviceGuard guard_(device);
    CU(guard_.err);
    aoed_gp* gp = new aoed_gp();
    struct Guard {  // a failure below must not leak the half-built handle (streams, events)
        aoed_gp* h;
        ~Guard() {
            if (h) aoed_gp_destroy(h);
        }
    } guard{gp};
    gp->device = device; gp->d = n_var; gp->m = n_obj; gp->nu = nu; gp->DP = DP;
    gp->DE = pick_de(n_var, DP);
    query_occupancy(nu, DP, gp->DE, gp->occ_xcov, gp->occ_gradsum);
    gp->trmm_variant = trmm_variant_from_env();
    gp->ssf = gp->trmm_variant == 0 ? PT_SS_PER_TILE : 1;
    CU(cudaStreamCreateWithFlags(&gp->streams[0], cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&gp->streams[1], cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&gp->d2h_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&gp->ev_h2d[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&gp->ev_done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&gp->ev_free[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreate(&gp->ev0));
    CU(cudaEventCreate(&gp->ev1));
    guard.h = nullptr;
    *out = gp;
    return AOED_OK;
}

int aoed_gp_destroy(aoed_gp_t* gp) {
    if (!gp) return AOED_OK;
    DeviceGuard guard_(gp->device);
    cudaDeviceSynchronize();
    free_model(gp);
    if (gp->ev0) cudaEventDestroy(gp->ev0);
    if (gp->ev1) cudaEventDestroy(gp->ev1);
    for (auto& s : gp->streams)
        if (s) cudaStreamDestroy(s);
    if (gp->d2h_stream) cudaStreamDestroy(gp->d2h_stream);
    for (int i = 0; i < 2; ++i) {
        if (gp->ev_h2d[i]) cudaEventDestroy(gp->ev_h2d[i]);
        if (gp->ev_done[i]) cudaEventDestroy(gp->ev_done[i]);
        if (gp->ev_free[i]) cudaEventDestroy(gp->ev_free[i]);
    }
    delete gp;
    return AOED_OK;
}

int aoed_gp_set_train(aoed_gp_t* gp, int n_train, const double* h_Xn, const double* h_Yn, const double* h_y_mean,
                      const double* h_y_scale) {
    if (!gp || !h_Xn || !h_Yn) return fail(AOED_ERR_INVALID, "null argument");
    if (n_train < 1) return fail(AOED_ERR_INVALID, "n_train must be positive");
    DeviceGuard guard_(gp->device);
    CU(guard_.err);
    CU(cudaDeviceSynchronize());
    const int N = n_train, d = gp->d, m = gp->m;
    const int prevN = gp->N;
    // MOBO refits every iteration with a few more rows: keep every allocation (model, fit scratch, workspaces, cuSOLVER
    // handles, TMA descriptors) while N stays in the same 128-row bucket
    const bool reuse = gp->cap_Npad == round_up(N, TM);
    if (!reuse) free_model(gp);
    gp->N = N;
    gp->Npad = round_up(N, TM);
    gp->Kpad = round_up(N, TK);
    gp->mTiles = gp->Npad / TM;
    gp->JS = 16;  // capacity of the training-index split (partial-sum buffers); the split per launch is choose_split()
    gp->jchunk = round_up((N + gp->JS - 1) / gp->JS, XC_JT);
    gp->has_train = true;
    gp->has_theta = false;
    gp->Xn.assign(h_Xn, h_Xn + size_t(N) * d);
    gp->Yn.assign(h_Yn, h_Yn + size_t(N) * m);
    gp->y_mean.assign(m, 0.0);
    gp->y_scale.assign(m, 1.0);
    if (h_y_mean) gp->y_mean.assign(h_y_mean, h_y_mean + m);
    if (h_y_scale) gp->y_scale.assign(h_y_scale, h_y_scale + m);
    gp->lml.assign(m, std::numeric_limits<double>::quiet_NaN());
    const size_t Np = gp->Npad;
    if (!reuse) {
        CU(cudaMalloc(&gp->dX, Np * d * sizeof(double)));
        CU(cudaMalloc(&gp->dXT, Np * d * sizeof(double)));
        CU(cudaMalloc(&gp->dY, size_t(m) * Np * sizeof(double)));
        CU(cudaMalloc(&gp->dTs, size_t(m) * Np * gp->DP * sizeof(double)));
        CU(cudaMalloc(&gp->dAlpha, size_t(m) * Np * sizeof(double)));
        CU(cudaMalloc(&gp->dEll, size_t(m) * gp->DP * sizeof(double)));
        CU(cudaMalloc(&gp->dModel, size_t(m) * sizeof(ModelDev)));
        CU(cudaMalloc(&gp->dLinv, size_t(m) * Np * Np * sizeof(double)));
        CU(cudaMalloc(&gp->dLinvT, size_t(m) * Np * Np * sizeof(double)));
        CU(cudaMalloc(&gp->dL, size_t(m) * Np * Np * sizeof(double)));
        CU(cudaMalloc(&gp->dTickets, 64 * sizeof(int)));  // ring for the compute stream
        if (gp->trmm_variant == 0) {
            CHK(make_map_2d(&gp->mapLinv, gp->dLinv, int64_t(m) * Np, Np, Np, TM));
            CHK(make_map_2d(&gp->mapLinvT, gp->dLinvT, int64_t(m) * Np, Np, Np, TM));
        }
        CU(cudaMalloc(&gp->dC1C2, size_t(m) * 2 * sizeof(double)));
        gp->cap_Npad = gp->Npad;
    } else if (prevN != N) {
        // slab rows in [N, Kpad) feed the triangular products as k-rows against zero columns of the factor: keep them
        // free of stale values from the previous training set
        for (auto& w : gp->ws) {
            if (w.cols == 0) continue;
            const size_t slab = size_t(m) * Np * size_t(w.cols) * sizeof(double);
            CU(cudaMemsetAsync(w.KsT, 0, slab, nullptr));
            CU(cudaMemsetAsync(w.Gr, 0, slab, nullptr));
            CU(cudaMemsetAsync(w.W, 0, slab, nullptr));
            CU(cudaMemsetAsync(w.V, 0, slab, nullptr));
        }
        CU(cudaDeviceSynchronize());
    }
    CU(cudaMemcpy(gp->dX, h_Xn, size_t(N) * d * sizeof(double), cudaMemcpyHostToDevice));
    {
        std::vector<double> xt(Np * d, 0.0);
        for (int i = 0; i < N; ++i)
            for (int k = 0; k < d; ++k) xt[size_t(k) * Np + i] = h_Xn[size_t(i) * d + k];
        CU(cudaMemcpy(gp->dXT, xt.data(), xt.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    std::vector<double> yt(size_t(m) * Np, 0.0);
    for (int o = 0; o < m; ++o)
        for (int i = 0; i < N; ++i) yt[size_t(o) * Np + i] = h_Yn[size_t(i) * m + o];
    CU(cudaMemcpy(gp->dY, yt.data(), yt.size() * sizeof(double), cudaMemcpyHostToDevice));
    return AOED_OK;
}

}  // extern "C"

namespace {

// Uploads Ts = X / l (padded), c1/c2 for ONE hyper-parameter vector into slot `slot` of the per-objective arrays.
int upload_theta_slot(aoed_gp* gp, int slot, const double* theta, cudaStream_t s) {
    const int d = gp->d, DP = gp->DP, N = gp->N;
    const size_t Np = gp->Npad;
    std::vector<double> ell(DP, 1.0);
    for (int k = 0; k < d; ++k) ell[k] = exp(theta[1 + k]);
    std::vector<double> ts(Np * DP, 0.0);
    for (int i = 0; i < N; ++i)
        for (int k = 0; k < d; ++k) ts[size_t(i) * DP + k] = gp->Xn[size_t(i) * d + k] / ell[k];
    double c12[2] = {exp(theta[0]), exp(theta[d + 1])};
    CU(cudaMemcpyAsync(gp->dTs + size_t(slot) * Np * DP, ts.data(), ts.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(gp->dEll + size_t(slot) * DP, ell.data(), DP * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(gp->dC1C2 + size_t(slot) * 2, c12, 2 * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));  // host vectors go out of scope
    return AOED_OK;
}

int ensure_batch(aoed_gp* gp, int n_prob) {
    if (gp->batchCap >= n_prob) return AOED_OK;
    if (gp->dThetaB) cudaFree(gp->dThetaB);
    if (gp->dResB) cudaFree(gp->dResB);
    if (gp->dObjB) cudaFree(gp->dObjB);
    const int cap = std::max(n_prob, 64);
    const size_t p = gp->d + 2;
    CU(cudaMalloc(&gp->dThetaB, size_t(cap) * p * sizeof(double)));
    CU(cudaMalloc(&gp->dResB, size_t(cap) * (p + 1) * sizeof(double)));
    CU(cudaMalloc(&gp->dObjB, size_t(cap) * sizeof(int)));
    gp->batchCap = cap;
    return AOED_OK;
}

// Scratch of the blocked factorisation for up to `want` problems side by side (fewer when the budget — AOED_FIT_SCRATCH_GB,
// default 16 GB, and half of the free device memory — says so; the callers loop over groups of gp->mats.cap problems).
int ensure_blocked(aoed_gp* gp, int want) {
    const int Npad = gp->Npad, DP = gp->DP;
    if (gp->plan.T * TM != Npad) CHK(blocked_plan_build(gp->plan, Npad / TM));
    // nothing to do when the scratch already serves `want` problems per group, or as many as the budget allowed last time
    // (cudaMemGetInfo takes a device-wide lock: not something to call once per L-BFGS round)
    if (gp->mats.Npad == Npad && gp->mats.cap >= std::min(std::max(want, 1), gp->blocked_cap_max)) return AOED_OK;
    const size_t per = 4 * size_t(Npad) * Npad * sizeof(double);
    const char* e = getenv("AOED_FIT_SCRATCH_GB");
    double budget = (e ? atof(e) : 16.0) * double(1ull << 30);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        const double held = double(gp->mats.cap) * double(per);  // what this handle already holds counts as available
        budget = std::min(budget, 0.5 * (double(free_b) + held));
    }
    const int cap_max = std::max(1, int(budget / double(per)));
    gp->blocked_cap_max = cap_max;
    const int cap = std::min(std::max(want, 1), cap_max);
    if (gp->mats.Npad == Npad && gp->mats.cap >= cap) return AOED_OK;
    blocked_mats_free(gp->mats);
    double** ptrs[] = {&gp->bTs, &gp->bC1C2, &gp->bZ, &gp->bAlpha, &gp->bPartial};
    for (auto pp : ptrs) {
        if (*pp) cudaFree(*pp);
        *pp = nullptr;
    }
    if (gp->bInfo) cudaFree(gp->bInfo);
    if (gp->bCounters) cudaFree(gp->bCounters);
    gp->bInfo = gp->bCounters = nullptr;
    CHK(blocked_mats_alloc(gp->mats, Npad, cap));
    const size_t nbp = size_t(Npad) / KM_T;
    CU(cudaMalloc(&gp->bTs, size_t(cap) * Npad * DP * sizeof(double)));
    CU(cudaMalloc(&gp->bC1C2, size_t(cap) * 2 * sizeof(double)));
    CU(cudaMalloc(&gp->bZ, size_t(cap) * Npad * sizeof(double)));
    CU(cudaMalloc(&gp->bAlpha, size_t(cap) * Npad * sizeof(double)));
    CU(cudaMalloc(&gp->bPartial, size_t(cap) * nbp * nbp * (DP + 2) * sizeof(double)));
    CU(cudaMalloc(&gp->bInfo, size_t(cap) * sizeof(int)));
    CU(cudaMalloc(&gp->bCounters, size_t(gp->plan.n_counters) * sizeof(int)));
    return AOED_OK;
}

// One group of G <= gp->mats.cap problems on stream s:  K(theta) -> L (mats.K), L^-1 (mats.X), L^-T (mats.XT),
// z = L^-1 y (bZ), alpha = L^-T z (`alpha_out`, stride Npad), bInfo[b] != 0 where K is not positive definite.
//   d_Ts [G][Npad][DP], d_c1c2 [G][2]: scaled inputs / kernel constants of the G problems; d_obj [G]: which y each uses.
int blocked_group(aoed_gp* gp, int G, const double* d_Ts, const double* d_c1c2, const int* d_obj, double* alpha_out,
                  cudaStream_t s) {
    const int N = gp->N, Npad = gp->Npad, DP = gp->DP;
    const int64_t NN = int64_t(Npad) * Npad;
    BlockedMats& mt = gp->mats;
    KmatArgs ka;
    ka.Ts = d_Ts; ka.c1c2 = d_c1c2; ka.K = mt.K; ka.N = N; ka.Npad = Npad; ka.DP = DP; ka.pad = 1;
    const int nbk = Npad / KM_T;
    launch_kmat(gp->nu, DP, ka, dim3(nbk, nbk, G), s);
    gp->n_launches++;
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(gp->bInfo, 0, size_t(G) * sizeof(int), s));
    CHK(blocked_chol_inverse(gp->plan, mt, G, gp->bInfo, gp->bCounters, sm_count(gp->device), s, &gp->n_launches));
    transpose_square_kernel<<<dim3(Npad / 32, Npad / 32, G), dim3(32, 8), 0, s>>>(mt.X, mt.XT, Npad);
    tri_matvec_kernel<false><<<dim3((N + 7) / 8, G), 256, 0, s>>>(mt.X, NN, gp->dY, Npad, d_obj, N, Npad, gp->bZ, Npad);
    tri_matvec_kernel<true><<<dim3((N + 7) / 8, G), 256, 0, s>>>(mt.XT, NN, gp->bZ, Npad, nullptr, N, Npad, alpha_out, Npad);
    gp->n_launches += 3;
    CU(cudaGetLastError());
    return AOED_OK;
}

}  // namespace

extern "C" {

int aoed_gp_set_theta(aoed_gp_t* gp, const double* h_theta) {
    if (!gp || !h_theta) return fail(AOED_ERR_INVALID, "null argument");
    if (!gp->has_train) return fail(AOED_ERR_INVALID, "set_train must be called before set_theta");
    DeviceGuard guard_(gp->device);
    CU(guard_.err);
    cudaStream_t s = gp->streams[0];
    const int N = gp->N, Npad = gp->Npad, m = gp->m, d = gp->d, p = d + 2, DP = gp->DP;
    const size_t NN = size_t(Npad) * Npad;
    gp->theta.assign(h_theta, h_theta + size_t(m) * p);
    gp->has_theta = false;
    std::vector<ModelDev> md(m);
    for (int o = 0; o < m; ++o) {
        const double* th = h_theta + size_t(o) * p;
        CHK(upload_theta_slot(gp, o, th, s));
        md[o].c1 = exp(th[0]);
        md[o].c2 = exp(th[d + 1]);
        md[o].y_mean = gp->y_mean[o];
        md[o].y_scale = gp->y_scale[o];
    }
    // all objectives side by side through the blocked factorisation (groups when the scratch budget is smaller than m)
    CHK(ensure_batch(gp, m));
    CHK(ensure_blocked(gp, m));
    std::vector<int> iota(m);
    for (int o = 0; o < m; ++o) iota[o] = o;
    CU(cudaMemcpyAsync(gp->dObjB, iota.data(), size_t(m) * sizeof(int), cudaMemcpyHostToDevice, s));
    std::vector<double> res(size_t(m) * (p + 1));
    std::vector<int> info(m, 0);
    const int cap = gp->mats.cap;
    for (int o0 = 0; o0 < m; o0 += cap) {
        const int G = std::min(cap, m - o0);
        double* alpha = gp->dAlpha + size_t(o0) * Npad;
        CHK(blocked_group(gp, G, gp->dTs + size_t(o0) * Npad * DP, gp->dC1C2 + size_t(o0) * 2, gp->dObjB + o0, alpha, s));
        // log marginal likelihood value at theta (_gpr.py:601-616)
        lml_value_kernel<<<G, 256, 0, s>>>(gp->bZ, gp->bZ, gp->mats.K, N, Npad, gp->bInfo, gp->dResB + size_t(o0) * (p + 1), p + 1);
        // cached state: L_ (lower), L^-1 and its transpose, zero outside the N x N lower triangle
        double* L = gp->dL + o0 * NN;
        double* Linv = gp->dLinv + o0 * NN;
        CU(cudaMemcpyAsync(L, gp->mats.K, size_t(G) * NN * sizeof(double), cudaMemcpyDeviceToDevice, s));
        CU(cudaMemcpyAsync(Linv, gp->mats.X, size_t(G) * NN * sizeof(double), cudaMemcpyDeviceToDevice, s));
        keep_lower_kernel<<<dim3((Npad + 127) / 128, Npad, G), 128, 0, s>>>(L, N, Npad);
        keep_lower_kernel<<<dim3((Npad + 127) / 128, Npad, G), 128, 0, s>>>(Linv, N, Npad);
        transpose_square_kernel<<<dim3(Npad / 32, Npad / 32, G), dim3(32, 8), 0, s>>>(Linv, gp->dLinvT + o0 * NN, Npad);
        gp->n_launches += 4;
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(info.data() + o0, gp->bInfo, size_t(G) * sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
    }
    for (int o = 0; o < m; ++o)
        if (info[o] != 0)
            return fail(AOED_ERR_NOT_PD, "kernel matrix of objective " + std::to_string(o) +
                                             " is not positive definite at the given theta (leading minor " +
                                             std::to_string(info[o]) + ")");
    CU(cudaMemcpyAsync(res.data(), gp->dResB, res.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(gp->dModel, md.data(), m * sizeof(ModelDev), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
    for (int o = 0; o < m; ++o) gp->lml[o] = res[size_t(o) * (p + 1)];
    gp->has_theta = true;
    return AOED_OK;
}

int aoed_gp_get_state(aoed_gp_t* gp, int obj, double* h_L, double* h_alpha, double* h_lml) {
    if (!gp || !gp->has_theta) return fail(AOED_ERR_INVALID, "model has no fitted state");
    if (obj < 0 || obj >= gp->m) return fail(AOED_ERR_INVALID, "objective index out of range");
    DeviceGuard guard_(gp->device);
    CU(guard_.err);
    const int N = gp->N, Npad = gp->Npad;
    if (h_L)
        CU(cudaMemcpy2D(h_L, size_t(N) * sizeof(double), gp->dL + size_t(obj) * Npad * Npad, size_t(Npad) * sizeof(double),
                        size_t(N) * sizeof(double), N, cudaMemcpyDeviceToHost));
    if (h_alpha)
        CU(cudaMemcpy(h_alpha, gp->dAlpha + size_t(obj) * Npad, size_t(N) * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_lml) *h_lml = gp->lml[obj];
    return AOED_OK;
}

int aoed_gp_lml_batch(aoed_gp_t* gp, int n_prob, const int32_t* h_obj, const double* h_theta, double* h_lml,
                      double* h_grad) {
    if (!gp || !h_obj || !h_theta || !h_lml) return fail(AOED_ERR_INVALID, "null argument");
    if (!gp->has_train) return fail(AOED_ERR_INVALID, "set_train must be called first");
    if (n_prob <= 0) return AOED_OK;
    static const bool fit_timing = getenv("AOED_FIT_TIMING") != nullptr;
    const auto t_call = std::chrono::steady_clock::now();
    DeviceGuard guard_(gp->device);
    CU(guard_.err);
    const int N = gp->N, Npad = gp->Npad, m = gp->m, d = gp->d, p = d + 2, DP = gp->DP;
    for (int b = 0; b < n_prob; ++b)
        if (h_obj[b] < 0 || h_obj[b] >= m) return fail(AOED_ERR_INVALID, "objective index out of range");
    CHK(ensure_batch(gp, n_prob));
    cudaStream_t s0 = gp->streams[0];
    CU(cudaMemcpyAsync(gp->dThetaB, h_theta, size_t(n_prob) * p * sizeof(double), cudaMemcpyHostToDevice, s0));
    CU(cudaMemcpyAsync(gp->dObjB, h_obj, size_t(n_prob) * sizeof(int), cudaMemcpyHostToDevice, s0));
    const bool want_grad = h_grad != nullptr;
    const char* force = getenv("AOED_FIT_PATH");  // "blocked" forces the large-N path (tests)
    static const bool force_v1 = getenv("AOED_FIT_SMALL") && strcmp(getenv("AOED_FIT_SMALL"), "v1") == 0;
    const bool blk_fits = !force_v1 && lml_blk_smem(N, DP) <= 227 * 1024;
    // One CTA per problem while the problem fits in shared memory.  Just above the blocked kernel's limit (N ~ 216) only
    // the rank-1 kernel fits (~1 ms per launch): worth it for many problems at once, but a handful of problems is served
    // faster by the multi-stream path (0.8 ms per round at N = 300).
    const bool few = n_prob <= 6 && !force_v1 && !(force && strcmp(force, "small") == 0);
    const bool small = N <= FS_MAX_N && !(force && strcmp(force, "blocked") == 0) && (blk_fits || !few);
    if (small) {
        // one CTA per problem, everything in shared memory
        LmlSmallArgs a;
        a.X = gp->dX; a.XT = gp->dXT; a.Y = gp->dY; a.theta = gp->dThetaB; a.obj = gp->dObjB; a.res = gp->dResB;
        a.N = N; a.Npad = Npad; a.d = d; a.want_grad = want_grad ? 1 : 0;
        a.clk = nullptr;
        static const bool stamp = getenv("AOED_FIT_STAMP") != nullptr;  // tuning aid: prints per-phase cycles of problem 0
        long long* dclk = nullptr;
        if (stamp) {
            CU(cudaMalloc(&dclk, size_t(n_prob) * 8 * sizeof(long long)));
            a.clk = dclk;
        }
        // blocked DMMA kernel when its panel buffer fits beside the packed triangle (N <= ~208), else the rank-1 kernel
        const bool blocked = blk_fits;
        if (blocked) CU(aoed::launch_lml_blk(gp->nu, DP, a, n_prob, s0));
        else CU(aoed::launch_lml_small(gp->nu, DP, a, n_prob, s0));
        if (stamp) {
            long long h[8];
            CU(cudaMemcpyAsync(h, dclk, sizeof(h), cudaMemcpyDeviceToHost, s0));
            CU(cudaStreamSynchronize(s0));
            fprintf(stderr, "lml_small%s N=%d cycles: kmat %lld chol %lld inv %lld alpha %lld grad %lld", blocked ? "(blocked)" : "", N,
                    h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4]);
            if (blocked) fprintf(stderr, " | diagonal blocks: factor %lld invert+store %lld", h[6], h[7]);
            fprintf(stderr, "\n");
            cudaFree(dclk);
        }
        gp->n_launches++;
        gp->n_fit_small += n_prob;
    } else {
        // blocked factorisation, all problems of a group side by side in every launch (blocked.cu)
        CHK(ensure_blocked(gp, n_prob));
        const int cap = gp->mats.cap;
        if (fit_timing) CU(cudaEventRecord(gp->ev0, s0));
        for (int g0 = 0; g0 < n_prob; g0 += cap) {
            const int G = std::min(cap, n_prob - g0);
            double* res = gp->dResB + size_t(g0) * (p + 1);
            theta_prep_kernel<<<dim3((Npad * DP + 255) / 256, G), 256, 0, s0>>>(gp->dX, gp->dThetaB + size_t(g0) * p, N, Npad, d,
                                                                                DP, gp->bTs, gp->bC1C2);
            gp->n_launches++;
            CHK(blocked_group(gp, G, gp->bTs, gp->bC1C2, gp->dObjB + g0, gp->bAlpha, s0));
            // lml = -1/2 z.z - sum log L_ii - N/2 log 2 pi with z = L^-1 y (the same number as y.alpha, better conditioned)
            lml_value_kernel<<<G, 256, 0, s0>>>(gp->bZ, gp->bZ, gp->mats.K, N, Npad, gp->bInfo, res, p + 1);
            gp->n_launches++;
            if (want_grad) {
                CHK(blocked_kinv(gp->plan, gp->mats, G, gp->bCounters, sm_count(gp->device), s0, &gp->n_launches));
                LmlGradArgs ga;
                ga.Ts = gp->bTs; ga.c1c2 = gp->bC1C2; ga.alpha = gp->bAlpha; ga.Kinv = gp->mats.P; ga.partial = gp->bPartial;
                const int nb = (N + KM_T - 1) / KM_T;  // blocks beyond nb x nb hold padding only
                ga.N = N; ga.Npad = Npad; ga.DP = DP; ga.nBlocksX = nb; ga.nBlocksY = nb;
                launch_lmlgrad(gp->nu, DP, ga, dim3(nb, nb, G), s0);
                lml_grad_reduce_kernel<<<dim3(DP + 2, G), 256, 0, s0>>>(gp->bPartial, nb * nb, DP, d, gp->bInfo, res);
                gp->n_launches += 2;
            }
            CU(cudaGetLastError());
        }
        gp->n_fit_blocked += n_prob;
    }
    std::vector<double> res(size_t(n_prob) * (p + 1));
    CU(cudaMemcpyAsync(res.data(), gp->dResB, res.size() * sizeof(double), cudaMemcpyDeviceToHost, s0));
    if (fit_timing && !small) CU(cudaEventRecord(gp->ev1, s0));
    CU(cudaStreamSynchronize(s0));
    if (fit_timing && !small) {  // AOED_FIT_TIMING=1: device time of the round (events) beside the host's wall time of the call
        float ms = 0.f;
        cudaEventElapsedTime(&ms, gp->ev0, gp->ev1);
        const double wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count();
        fprintf(stderr, "[fit timing] problems %d device %.2f ms call %.2f ms\n", n_prob, double(ms), wall);
    }
    for (int b = 0; b < n_prob; ++b) {
        h_lml[b] = res[size_t(b) * (p + 1)];
        if (want_grad)
            for (int k = 0; k < p; ++k) h_grad[size_t(b) * p + k] = res[size_t(b) * (p + 1) + 1 + k];
    }
    return AOED_OK;
}

static int eval_common(aoed_gp_t* gp, int64_t n_rows, uint32_t flags, int acq_kind, const double* h_acq_param,
                       bool& want_std, bool& want_grad, AcqDev& acq) {
    if (!gp) return fail(AOED_ERR_INVALID, "null handle");
    if (!gp->has_theta) return fail(AOED_ERR_INVALID, "surrogate model is not fitted yet");
    if (n_rows < 0) return fail(AOED_ERR_INVALID, "negative row count");
    want_std = (flags & AOED_EVAL_STD) != 0 || acq_kind >= AOED_ACQ_UCB;
    want_grad = (flags & AOED_EVAL_GRAD) != 0;
    CHK(make_acq(gp, acq_kind, flags, h_acq_param, acq));
    return AOED_OK;
}

int aoed_gp_eval_device(aoed_gp_t* gp, int64_t n_rows, const double* d_Xn, uint32_t flags, int acq_kind,
                        const double* h_acq_param, double* d_F, double* d_S, double* d_dF, double* d_dS,
                        double* d_hF, double* d_hS, double* d_A, double* d_dA, double* d_hA, void* stream) {
    bool want_std, want_grad;
    AcqDev acq;
    CHK(eval_common(gp, n_rows, flags, acq_kind, h_acq_param, want_std, want_grad, acq));
    if (n_rows == 0) return AOED_OK;
    if (!d_Xn) return fail(AOED_ERR_INVALID, "null input");
    DeviceGuard guard_(gp->device);
    CU(guard_.err);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bool hess = (flags & AOED_EVAL_HESS) != 0;
    const int m = gp->m, d = gp->d;
    Workspace& w = gp->ws[0];
    int64_t step;
    if (hess) {
        step = hess_rows_cap(gp, n_rows);
        CHK(ensure_ws_hess(gp, w, step, false));
    } else {
        CHK(ensure_ws(gp, w, std::min<int64_t>(chunk_cap(true), round_up64(n_rows, 128)), false));
        step = w.cols;
    }
    for (int64_t r0 = 0; r0 < n_rows; r0 += step) {
        const int64_t rows = std::min<int64_t>(step, n_rows - r0);
        EvalOut out;
        out.F = d_F ? d_F + r0 * m : nullptr;
        out.S = d_S ? d_S + r0 * m : nullptr;
        out.A = d_A ? d_A + r0 * m : nullptr;
        out.dF = d_dF ? d_dF + r0 * m * d : nullptr;
        out.dS = d_dS ? d_dS + r0 * m * d : nullptr;
        out.dA = d_dA ? d_dA + r0 * m * d : nullptr;
        if (hess) {
            HessOut ho;
            ho.hF = d_hF ? d_hF + r0 * m * d * d : nullptr;
            ho.hS = d_hS ? d_hS + r0 * m * d * d : nullptr;
            ho.hA = d_hA ? d_hA + r0 * m * d * d : nullptr;
            CHK(eval_hess_chunk(gp, w, d_Xn + r0 * d, rows, want_std, acq, out, ho, s));
        } else {
            CHK(eval_chunk(gp, w, d_Xn + r0 * d, rows, want_std, want_grad, acq, out, s));
        }
    }
    return AOED_OK;
}

int aoed_gp_eval(aoed_gp_t* gp, int64_t n_rows, const double* h_Xn, uint32_t flags, int acq_kind,
                 const double* h_acq_param, double* h_F, double* h_S, double* h_dF, double* h_dS, double* h_hF,
                 double* h_hS, double* h_A, double* h_dA, double* h_hA) {
    bool want_std, want_grad;
    AcqDev acq;
    CHK(eval_common(gp, n_rows, flags, acq_kind, h_acq_param, want_std, want_grad, acq));
    if (n_rows == 0) return AOED_OK;
    if (!h_Xn) return fail(AOED_ERR_INVALID, "null input");
    DeviceGuard guard_(gp->device);
    CU(guard_.err);
    const int m = gp->m, d = gp->d;
    const bool hess = (flags & AOED_EVAL_HESS) != 0;
    int64_t cap;
    int nslots;
    if (hess) {
        cap = hess_rows_cap(gp, n_rows);
        nslots = (n_rows > cap) ? 2 : 1;
        for (int i = 0; i < nslots; ++i) CHK(ensure_ws_hess(gp, gp->ws[i], cap, true));
    } else {
        cap = std::min<int64_t>(chunk_cap(), round_up64(n_rows, 128));
        nslots = (n_rows > cap) ? 2 : 1;
        for (int i = 0; i < nslots; ++i) CHK(ensure_ws(gp, gp->ws[i], cap, true));
    }
    // Three in-order queues: H2D copies, compute, D2H copies, chained by events per staging slot.  The kernels of
    // consecutive chunks therefore run strictly back to back on ONE stream (exactly like the device-resident API) while
    // the PCIe traffic of the previous / next chunk overlaps them; two streams of kernels would interleave the persistent
    // triangular products of two chunks and lose ~4 % (measured).
    cudaStream_t sc = gp->streams[0], sh = gp->streams[1], sd = gp->d2h_stream;
    CU(cudaStreamSynchronize(sc));  // earlier device-API work on the compute stream's workspaces
    struct DrainGuard {  // an error return below must not leave copies into the caller's (pooled, recyclable) buffers in flight
        cudaStream_t a, b, c;
        ~DrainGuard() {
            cudaStreamSynchronize(a);
            cudaStreamSynchronize(b);
            cudaStreamSynchronize(c);
        }
    } drain{sh, sc, sd};
    int slot = 0;
    bool used[2] = {false, false};
    for (int64_t r0 = 0; r0 < n_rows; r0 += cap, slot = (slot + 1) % nslots) {
        Workspace& w = gp->ws[slot];
        const int64_t rows = std::min<int64_t>(cap, n_rows - r0);
        // the slot's previous chunk must have drained (its D2H read the staging buffers) before they are reused
        if (used[slot]) CU(cudaEventSynchronize(gp->ev_free[slot]));
        used[slot] = true;
        memcpy(w.hX, h_Xn + r0 * d, size_t(rows) * d * sizeof(double));
        CU(cudaMemcpyAsync(w.Xc, w.hX, size_t(rows) * d * sizeof(double), cudaMemcpyHostToDevice, sh));
        CU(cudaEventRecord(gp->ev_h2d[slot], sh));
        CU(cudaStreamWaitEvent(sc, gp->ev_h2d[slot], 0));
        const bool grad_out = want_grad;  // gradients are computed internally on the Hessian path either way
        EvalOut out;
        out.F = h_F ? w.oF : nullptr;
        out.S = (h_S && want_std) ? w.oS : nullptr;
        out.A = h_A ? w.oA : nullptr;
        out.dF = (h_dF && grad_out) ? w.odF : nullptr;
        out.dS = (h_dS && grad_out && want_std) ? w.odS : nullptr;
        out.dA = (h_dA && grad_out) ? w.odA : nullptr;
        HessOut ho{nullptr, nullptr, nullptr};
        if (hess) {
            ho.hF = h_hF ? w.ohF : nullptr;
            ho.hS = (h_hS && want_std) ? w.ohS : nullptr;
            ho.hA = (h_hA && acq.kind != AOED_ACQ_NONE) ? w.ohA : nullptr;
            CHK(eval_hess_chunk(gp, w, w.Xc, rows, want_std, acq, out, ho, sc));
        } else {
            CHK(eval_chunk(gp, w, w.Xc, rows, want_std, want_grad, acq, out, sc));
        }
        CU(cudaEventRecord(gp->ev_done[slot], sc));
        CU(cudaStreamWaitEvent(sd, gp->ev_done[slot], 0));
        const size_t b1 = size_t(rows) * m * sizeof(double), b2 = b1 * d, b3 = b2 * d;
        if (out.F) CU(cudaMemcpyAsync(h_F + r0 * m, w.oF, b1, cudaMemcpyDeviceToHost, sd));
        if (out.S) CU(cudaMemcpyAsync(h_S + r0 * m, w.oS, b1, cudaMemcpyDeviceToHost, sd));
        if (out.A) CU(cudaMemcpyAsync(h_A + r0 * m, w.oA, b1, cudaMemcpyDeviceToHost, sd));
        if (out.dF) CU(cudaMemcpyAsync(h_dF + r0 * m * d, w.odF, b2, cudaMemcpyDeviceToHost, sd));
        if (out.dS) CU(cudaMemcpyAsync(h_dS + r0 * m * d, w.odS, b2, cudaMemcpyDeviceToHost, sd));
        if (out.dA) CU(cudaMemcpyAsync(h_dA + r0 * m * d, w.odA, b2, cudaMemcpyDeviceToHost, sd));
        if (ho.hF) CU(cudaMemcpyAsync(h_hF + r0 * m * d * d, w.ohF, b3, cudaMemcpyDeviceToHost, sd));
        if (ho.hS) CU(cudaMemcpyAsync(h_hS + r0 * m * d * d, w.ohS, b3, cudaMemcpyDeviceToHost, sd));
        if (ho.hA) CU(cudaMemcpyAsync(h_hA + r0 * m * d * d, w.ohA, b3, cudaMemcpyDeviceToHost, sd));
        CU(cudaEventRecord(gp->ev_free[slot], sd));
    }
    CU(cudaStreamSynchronize(sd));
    CU(cudaStreamSynchronize(sc));
    return AOED_OK;
}

int aoed_gp_set_profiling(aoed_gp_t* gp, int on) {
    if (!gp) return fail(AOED_ERR_INVALID, "null handle");
    gp->profiling = on != 0;
    return AOED_OK;
}

int aoed_gp_get_counters(aoed_gp_t* gp, int64_t* n_launches, double* trmm_ms, int64_t* trmm_launches,
                         double* trmm_flops) {
    if (!gp) return fail(AOED_ERR_INVALID, "null handle");
    if (n_launches) *n_launches = gp->n_launches;
    if (trmm_ms) *trmm_ms = gp->trmm_ms;
    if (trmm_launches) *trmm_launches = gp->trmm_launches;
    if (trmm_flops) *trmm_flops = gp->trmm_flops;
    return AOED_OK;
}

int aoed_gp_get_fit_counters(aoed_gp_t* gp, int64_t* n_small, int64_t* n_blocked) {
    if (!gp) return fail(AOED_ERR_INVALID, "null handle");
    if (n_small) *n_small = gp->n_fit_small;
    if (n_blocked) *n_blocked = gp->n_fit_blocked;
    return AOED_OK;
}

int aoed_gp_reset_counters(aoed_gp_t* gp) {
    if (!gp) return fail(AOED_ERR_INVALID, "null handle");
    gp->n_launches = gp->trmm_launches = 0;
    gp->trmm_ms = gp->trmm_flops = 0.0;
    return AOED_OK;
}

}  // extern "C"
