// Multi-objective selection kernels (sm_100a): Pareto mask / rank and greedy hypervolume improvement.
// Contracts in include/aoed_b200.h.  Replaces autooed/utils/pareto.py:27-62 and
// autooed/mobo/selection/hvi.py:16-47 (+ the pymoo hypervolume it calls).
//
// Hypervolume improvement is computed DIRECTLY as the exclusive volume of box(y, ref) not dominated by the
// front, by a dimension sweep: slabs along the last objective, recursing down to a 2-D staircase sweep over
// the front sorted by its first objective.  A dominated candidate yields exactly 0, as in the reference where
// pymoo's non-dominated pre-filter makes HV(P u {y}) - HV(P) vanish.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/aoed_b200.h"
#include "common.cuh"
#include "rank.cuh"

extern "C" int aoed_detail_fail(int code, const char* msg);  // defined in gp.cu (shared last-error slot)

namespace {

using aoed::RS_MAX_N;

int mfail(int code, const std::string& msg) { return aoed_detail_fail(code, msg.c_str()); }
#define MCU(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return mfail(AOED_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + \
                                            __FILE__ + ":" + std::to_string(__LINE__));             \
    } while (0)

constexpr int MOO_MAX_M = 8;

// ---- Pareto ------------------------------------------------------------------------------------------
// YT: [m][n] objective-major.  alive == nullptr -> all points take part.  dominated[i] = 1 iff some alive j
// has all(Y_j <= Y_i) and any(Y_j < Y_i)   (utils/pareto.py:39).
template <int M>
__global__ void __launch_bounds__(256) dominated_kernel(const double* __restrict__ YT, int64_t n, int64_t row0,
                                                        int64_t nrows, const uint8_t* __restrict__ alive,
                                                        uint8_t* __restrict__ dominated) {
    __shared__ double tile[M][256];
    __shared__ uint8_t talive[256];
    const int64_t li = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;  // row of this shard
    const int64_t i = row0 + li;
    double yi[M];
#pragma unroll
    for (int k = 0; k < M; ++k) yi[k] = (li < nrows) ? YT[k * n + i] : 0.0;
    bool dom = false;
    for (int64_t j0 = 0; j0 < n; j0 += 256) {
        const int64_t j = j0 + threadIdx.x;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < M; ++k) tile[k][threadIdx.x] = (j < n) ? YT[k * n + j] : DBL_MAX;
        talive[threadIdx.x] = (j < n) ? (alive ? alive[j] : 1) : 0;
        __syncthreads();
        const int lim = (n - j0 < 256) ? int(n - j0) : 256;
        for (int t = 0; t < lim; ++t) {
            bool le = true, lt = false;
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const double v = tile[k][t];
                le = le && (v <= yi[k]);
                lt = lt || (v < yi[k]);
            }
            dom = dom || (le && lt && talive[t]);
        }
    }
    if (li < nrows) dominated[li] = dom ? 1 : 0;
}

template <int M>
void launch_dominated_m(const double* YT, int64_t n, int64_t row0, int64_t nrows, const uint8_t* alive, uint8_t* dom,
                        cudaStream_t s) {
    dominated_kernel<M><<<unsigned((nrows + 255) / 256), 256, 0, s>>>(YT, n, row0, nrows, alive, dom);
}

int launch_dominated(int m, const double* YT, int64_t n, const uint8_t* alive, uint8_t* dom, cudaStream_t s,
                     int64_t row0 = 0, int64_t nrows = -1) {
    if (nrows < 0) nrows = n;
    if (nrows == 0) return AOED_OK;
    switch (m) {
        case 1: launch_dominated_m<1>(YT, n, row0, nrows, alive, dom, s); break;
        case 2: launch_dominated_m<2>(YT, n, row0, nrows, alive, dom, s); break;
        case 3: launch_dominated_m<3>(YT, n, row0, nrows, alive, dom, s); break;
        case 4: launch_dominated_m<4>(YT, n, row0, nrows, alive, dom, s); break;
        case 5: launch_dominated_m<5>(YT, n, row0, nrows, alive, dom, s); break;
        case 6: launch_dominated_m<6>(YT, n, row0, nrows, alive, dom, s); break;
        case 7: launch_dominated_m<7>(YT, n, row0, nrows, alive, dom, s); break;
        case 8: launch_dominated_m<8>(YT, n, row0, nrows, alive, dom, s); break;
        default: return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    }
    MCU(cudaGetLastError());
    return AOED_OK;
}

// rank pass: alive & not dominated -> rank = r, alive = 0; counts how many remain alive
__global__ void rank_update_kernel(int64_t n, const uint8_t* dominated, uint8_t* alive, int32_t* rank, int r,
                                   unsigned long long* remaining) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n || !alive[i]) return;
    if (!dominated[i]) {
        rank[i] = r;
        alive[i] = 0;
    } else {
        atomicAdd(remaining, 1ull);
    }
}

int launch_rank_small(int m, const double* YT, int n, int32_t* rank, int* cnt_scratch, cudaStream_t s) {
#define RS_CASE(MM)                                                                       \
    case MM:                                                                              \
        MCU(aoed::launch_rank_small_m<MM>(YT, n, rank, cnt_scratch, /*need=*/n, s));      \
        break;
    switch (m) {
        RS_CASE(1) RS_CASE(2) RS_CASE(3) RS_CASE(4) RS_CASE(5) RS_CASE(6) RS_CASE(7) RS_CASE(8)
        default: return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    }
#undef RS_CASE
    return AOED_OK;
}

// ---- exclusive hypervolume -----------------------------------------------------------------------------
// front: [k][M] sorted ascending by objective 0 (shared memory).  lev[] holds, for objectives >= D, the slab
// level: a front point is active iff p[t] <= lev[t] for all t in [D, M).
template <int M>
__device__ __forceinline__ bool is_active(const double* p, const double* lev, int D) {
    bool a = true;
#pragma unroll
    for (int t = 2; t < M; ++t)
        if (t >= D) a = a && (p[t] <= lev[t]);
    return a;
}

template <int M>
__device__ double excl2d(const double* y, const double* front, int k, const double* ref, const double* lev) {
    const double a = y[0], b = y[1];
    double hmin = ref[1], xprev = a, area = 0.0;
    for (int i = 0; i < k; ++i) {
        const double* p = front + i * M;
        if (!is_active<M>(p, lev, 2)) continue;
        const double px = p[0];
        if (px >= ref[0]) break;
        if (px > a) {
            const double h = hmin - b;
            if (h <= 0.0) return area;
            area = fma(px - xprev, h, area);
            xprev = px;
        }
        hmin = fmin(hmin, p[1]);
    }
    const double h = hmin - b;
    if (h > 0.0) area = fma(ref[0] - xprev, h, area);
    return area;
}

template <int M, int D>
struct Excl {
    // exclusive volume in objectives [0, D) of y, slabbing along objective D-1
    static __device__ double run(const double* y, const double* front, int k, const double* ref, double* lev) {
        double vol = 0.0;
        double cur = y[D - 1];
        const double top = ref[D - 1];
        while (cur < top) {
            // next slab boundary: smallest front coordinate above cur (among points active in higher dims)
            double nxt = top;
            for (int i = 0; i < k; ++i) {
                const double* p = front + i * M;
                const double v = p[D - 1];
                if (v > cur && v < nxt && is_active<M>(p, lev, D)) nxt = v;
            }
            lev[D - 1] = cur;
            const double sub = Excl<M, D - 1>::run(y, front, k, ref, lev);
            if (sub <= 0.0) break;  // the cross-section only shrinks as the level rises
            vol = fma(nxt - cur, sub, vol);
            cur = nxt;
        }
        return vol;
    }
};
template <int M>
struct Excl<M, 2> {
    static __device__ double run(const double* y, const double* front, int k, const double* ref, double* lev) {
        return excl2d<M>(y, front, k, ref, lev);
    }
};
template <int M>
struct Excl<M, 1> {
    static __device__ double run(const double* y, const double*, int, const double* ref, double*) { return ref[0] - y[0]; }
};

// One thread per candidate.  contrib[i] = exclusive hypervolume (0 for excluded / out-of-box candidates).
template <int M>
__global__ void __launch_bounds__(128) hvc_kernel(const double* __restrict__ YcT, int64_t n, const double* __restrict__ front,
                                                  int k, const double* __restrict__ ref, const uint8_t* __restrict__ excluded,
                                                  double* __restrict__ contrib) {
    extern __shared__ double sfront[];
    __shared__ double sref[M];
    for (int i = threadIdx.x; i < k * M; i += blockDim.x) sfront[i] = front[i];
    if (threadIdx.x < M) sref[threadIdx.x] = ref[threadIdx.x];
    __syncthreads();
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    double y[M], lev[M];
    bool inside = true;
#pragma unroll
    for (int t = 0; t < M; ++t) {
        y[t] = YcT[t * n + i];
        lev[t] = y[t];
        inside = inside && (y[t] <= sref[t]);
    }
    double v = 0.0;
    if (inside && !(excluded && excluded[i])) {
        if (M == 1) {
            double best = sref[0];
            for (int j = 0; j < k; ++j) best = fmin(best, sfront[j]);
            v = fmax(0.0, best - y[0]);
        } else {
            v = Excl<M, M>::run(y, sfront, k, sref, lev);
        }
    }
    contrib[i] = v;
}

int launch_hvc(int m, const double* YcT, int64_t n, const double* front, int k, const double* ref,
               const uint8_t* excluded, double* contrib, cudaStream_t s) {
    const unsigned grid = unsigned((n + 127) / 128);
#define HVC_CASE(MM)                                                                                            \
    case MM: {                                                                                                  \
        const size_t sm = size_t(std::max(k, 1)) * MM * sizeof(double);                                         \
        if (sm > 48 * 1024) MCU(cudaFuncSetAttribute(hvc_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sm))); \
        hvc_kernel<MM><<<grid, 128, sm, s>>>(YcT, n, front, k, ref, excluded, contrib);                         \
        break;                                                                                                  \
    }
    switch (m) {
        HVC_CASE(1)
        HVC_CASE(2)
        HVC_CASE(3)
        HVC_CASE(4)
        HVC_CASE(5)
        HVC_CASE(6)
        default: return mfail(AOED_ERR_UNSUPPORTED, "hypervolume improvement supports n_obj <= 6");
    }
#undef HVC_CASE
    MCU(cudaGetLastError());
    return AOED_OK;
}

// block-level arg-max with lowest-index tie-break (hvi.py:35-37: strict '>' keeps the first maximum)
__global__ void __launch_bounds__(256) argmax_stage1(const double* __restrict__ v, int64_t n, double* bval, int64_t* bidx) {
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    double best = -1.0;
    int64_t bi = -1;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const double x = v[i];
        if (x > best) {  // ascending i per thread: strict '>' keeps the lowest index
            best = x;
            bi = i;
        }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            const double o = sv[threadIdx.x + off];
            const int64_t oi = si[threadIdx.x + off];
            if (oi >= 0 && (o > sv[threadIdx.x] || (o == sv[threadIdx.x] && (si[threadIdx.x] < 0 || oi < si[threadIdx.x])))) {
                sv[threadIdx.x] = o;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bval[blockIdx.x] = sv[0];
        bidx[blockIdx.x] = si[0];
    }
}

struct Buf {
    void* p = nullptr;
    ~Buf() {
        if (p) cudaFree(p);
    }
    template <typename T>
    T* as() { return static_cast<T*>(p); }
};

std::vector<double> to_objective_major(const double* Y, int64_t n, int m) {
    std::vector<double> t(size_t(n) * m);
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < m; ++k) t[size_t(k) * n + i] = Y[i * m + k];
    return t;
}

std::vector<double> sort_front(const double* front, int64_t k, int m) {
    std::vector<int64_t> idx(k);
    std::iota(idx.begin(), idx.end(), 0);
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return front[a * m] < front[b * m]; });
    std::vector<double> out(size_t(k) * m);
    for (int64_t i = 0; i < k; ++i)
        for (int t = 0; t < m; ++t) out[size_t(i) * m + t] = front[idx[i] * m + t];
    return out;
}

}  // namespace

extern "C" {

int aoed_pareto_mask(int device, int64_t n, int m, const double* h_Y, uint8_t* h_mask) {
    return aoed_pareto_mask_rows(device, n, m, h_Y, 0, n, h_mask);
}

int aoed_pareto_mask_rows(int device, int64_t n, int m, const double* h_Y, int64_t row_lo, int64_t row_hi,
                          uint8_t* h_mask) {
    if (n < 0 || m < 1 || row_lo < 0 || row_hi < row_lo || row_hi > n || (n > 0 && !h_Y) || (row_hi > row_lo && !h_mask))
        return mfail(AOED_ERR_INVALID, "bad argument");
    const int64_t nrows = row_hi - row_lo;
    if (nrows == 0) return AOED_OK;
    if (m > MOO_MAX_M) return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return mfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    MCU(guard_.err);
    std::vector<double> yt = to_objective_major(h_Y, n, m);
    Buf dY, dDom;
    MCU(cudaMalloc(&dY.p, yt.size() * sizeof(double)));
    MCU(cudaMalloc(&dDom.p, size_t(nrows)));
    MCU(cudaMemcpy(dY.p, yt.data(), yt.size() * sizeof(double), cudaMemcpyHostToDevice));
    int rc = launch_dominated(m, dY.as<double>(), n, nullptr, dDom.as<uint8_t>(), nullptr, row_lo, nrows);
    if (rc != AOED_OK) return rc;
    MCU(cudaMemcpy(h_mask, dDom.p, size_t(nrows), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < nrows; ++i) h_mask[i] = h_mask[i] ? 0 : 1;
    return AOED_OK;
}

int aoed_pareto_rank(int device, int64_t n, int m, const double* h_Y, int32_t* h_rank) {
    if (n < 0 || m < 1 || (n > 0 && (!h_Y || !h_rank))) return mfail(AOED_ERR_INVALID, "bad argument");
    if (n == 0) return AOED_OK;
    if (m > MOO_MAX_M) return mfail(AOED_ERR_UNSUPPORTED, "n_obj > 8 is not supported");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return mfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    MCU(guard_.err);
    std::vector<double> yt = to_objective_major(h_Y, n, m);
    const char* force = getenv("AOED_RANK_PATH");  // "multi" forces the multi-launch path (tests)
    if (!(force && strcmp(force, "multi") == 0) && n <= RS_MAX_N &&
        size_t(m) * n * sizeof(double) + 3 * size_t(n) * sizeof(int) <= 200 * 1024) {
        // one launch, one allocation: the call is latency bound (it runs once per NSGA-II generation)
        Buf d;
        const size_t ybytes = yt.size() * sizeof(double);
        MCU(cudaMalloc(&d.p, ybytes + (1 + aoed::DC_SPLIT) * size_t(n) * sizeof(int32_t)));
        MCU(cudaMemcpy(d.p, yt.data(), ybytes, cudaMemcpyHostToDevice));
        int32_t* drank = reinterpret_cast<int32_t*>(static_cast<char*>(d.p) + ybytes);
        int rc = launch_rank_small(m, d.as<double>(), int(n), drank, drank + n, nullptr);
        if (rc != AOED_OK) return rc;
        MCU(cudaMemcpy(h_rank, drank, size_t(n) * sizeof(int32_t), cudaMemcpyDeviceToHost));
        return AOED_OK;
    }
    Buf dY, dDom, dAlive, dRank, dRem;
    MCU(cudaMalloc(&dY.p, yt.size() * sizeof(double)));
    MCU(cudaMalloc(&dDom.p, size_t(n)));
    MCU(cudaMalloc(&dAlive.p, size_t(n)));
    MCU(cudaMalloc(&dRank.p, size_t(n) * sizeof(int32_t)));
    MCU(cudaMalloc(&dRem.p, sizeof(unsigned long long)));
    MCU(cudaMemcpy(dY.p, yt.data(), yt.size() * sizeof(double), cudaMemcpyHostToDevice));
    MCU(cudaMemset(dAlive.p, 1, size_t(n)));
    MCU(cudaMemset(dRank.p, 0xff, size_t(n) * sizeof(int32_t)));
    for (int r = 0;; ++r) {
        int rc = launch_dominated(m, dY.as<double>(), n, dAlive.as<uint8_t>(), dDom.as<uint8_t>(), nullptr);
        if (rc != AOED_OK) return rc;
        MCU(cudaMemset(dRem.p, 0, sizeof(unsigned long long)));
        rank_update_kernel<<<unsigned((n + 255) / 256), 256>>>(n, dDom.as<uint8_t>(), dAlive.as<uint8_t>(),
                                                               dRank.as<int32_t>(), r, dRem.as<unsigned long long>());
        MCU(cudaGetLastError());
        unsigned long long rem = 0;
        MCU(cudaMemcpy(&rem, dRem.p, sizeof(rem), cudaMemcpyDeviceToHost));
        if (rem == 0) break;
        if (r > n) return mfail(AOED_ERR_CUDA, "pareto_rank did not terminate");
    }
    MCU(cudaMemcpy(h_rank, dRank.p, size_t(n) * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return AOED_OK;
}

int aoed_hv_contributions(int device, int64_t n_cand, int m, const double* h_Yc, int64_t n_front,
                          const double* h_front, const double* h_ref, double* h_contrib) {
    if (n_cand < 0 || m < 1 || n_front < 0 || !h_ref || (n_cand > 0 && (!h_Yc || !h_contrib)) || (n_front > 0 && !h_front))
        return mfail(AOED_ERR_INVALID, "bad argument");
    if (n_cand == 0) return AOED_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return mfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    MCU(guard_.err);
    std::vector<double> yt = to_objective_major(h_Yc, n_cand, m);
    std::vector<double> fs = sort_front(h_front, n_front, m);
    Buf dY, dF, dR, dC;
    MCU(cudaMalloc(&dY.p, yt.size() * sizeof(double)));
    MCU(cudaMalloc(&dF.p, std::max<size_t>(fs.size(), 1) * sizeof(double)));
    MCU(cudaMalloc(&dR.p, m * sizeof(double)));
    MCU(cudaMalloc(&dC.p, size_t(n_cand) * sizeof(double)));
    MCU(cudaMemcpy(dY.p, yt.data(), yt.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (!fs.empty()) MCU(cudaMemcpy(dF.p, fs.data(), fs.size() * sizeof(double), cudaMemcpyHostToDevice));
    MCU(cudaMemcpy(dR.p, h_ref, m * sizeof(double), cudaMemcpyHostToDevice));
    int rc = launch_hvc(m, dY.as<double>(), n_cand, dF.as<double>(), int(n_front), dR.as<double>(), nullptr,
                        dC.as<double>(), nullptr);
    if (rc != AOED_OK) return rc;
    MCU(cudaMemcpy(h_contrib, dC.p, size_t(n_cand) * sizeof(double), cudaMemcpyDeviceToHost));
    return AOED_OK;
}

int aoed_hvi_select(int device, int64_t n_cand, int m, const double* h_Yc, int64_t n_front, const double* h_front,
                    const double* h_ref, int batch, const uint8_t* h_excluded, int64_t* h_idx, double* h_contrib) {
    if (n_cand < 0 || m < 1 || n_front < 0 || batch < 0 || !h_ref || !h_idx || (n_cand > 0 && !h_Yc) ||
        (n_front > 0 && !h_front))
        return mfail(AOED_ERR_INVALID, "bad argument");
    for (int b = 0; b < batch; ++b) {
        h_idx[b] = -1;
        if (h_contrib) h_contrib[b] = 0.0;
    }
    if (n_cand == 0 || batch == 0) return AOED_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return mfail(AOED_ERR_NO_DEVICE, "no CUDA device");
    aoed::DeviceGuard guard_(device);
    MCU(guard_.err);
    std::vector<double> yt = to_objective_major(h_Yc, n_cand, m);
    std::vector<double> front(h_front, h_front + size_t(n_front) * m);
    std::vector<uint8_t> excl(size_t(n_cand), 0);
    if (h_excluded) excl.assign(h_excluded, h_excluded + n_cand);
    const int nblk = int(std::min<int64_t>(1024, (n_cand + 255) / 256));
    Buf dY, dF, dR, dC, dE, dBV, dBI;
    const size_t front_cap = size_t(n_front + batch) * m;
    MCU(cudaMalloc(&dY.p, yt.size() * sizeof(double)));
    MCU(cudaMalloc(&dF.p, std::max<size_t>(front_cap, 1) * sizeof(double)));
    MCU(cudaMalloc(&dR.p, m * sizeof(double)));
    MCU(cudaMalloc(&dC.p, size_t(n_cand) * sizeof(double)));
    MCU(cudaMalloc(&dE.p, size_t(n_cand)));
    MCU(cudaMalloc(&dBV.p, nblk * sizeof(double)));
    MCU(cudaMalloc(&dBI.p, nblk * sizeof(int64_t)));
    MCU(cudaMemcpy(dY.p, yt.data(), yt.size() * sizeof(double), cudaMemcpyHostToDevice));
    MCU(cudaMemcpy(dR.p, h_ref, m * sizeof(double), cudaMemcpyHostToDevice));
    MCU(cudaMemcpy(dE.p, excl.data(), size_t(n_cand), cudaMemcpyHostToDevice));
    std::vector<double> bv(nblk);
    std::vector<int64_t> bi(nblk);
    for (int b = 0; b < batch; ++b) {
        const int64_t k = int64_t(front.size() / m);
        std::vector<double> fs = sort_front(front.data(), k, m);
        if (!fs.empty()) MCU(cudaMemcpy(dF.p, fs.data(), fs.size() * sizeof(double), cudaMemcpyHostToDevice));
        int rc = launch_hvc(m, dY.as<double>(), n_cand, dF.as<double>(), int(k), dR.as<double>(), dE.as<uint8_t>(),
                            dC.as<double>(), nullptr);
        if (rc != AOED_OK) return rc;
        argmax_stage1<<<nblk, 256>>>(dC.as<double>(), n_cand, dBV.as<double>(), dBI.as<int64_t>());
        MCU(cudaGetLastError());
        MCU(cudaMemcpy(bv.data(), dBV.p, nblk * sizeof(double), cudaMemcpyDeviceToHost));
        MCU(cudaMemcpy(bi.data(), dBI.p, nblk * sizeof(int64_t), cudaMemcpyDeviceToHost));
        double best = 0.0;
        int64_t best_i = -1;
        for (int t = 0; t < nblk; ++t) {
            if (bi[t] < 0) continue;
            if (bv[t] > best || (bv[t] == best && best_i >= 0 && bi[t] < best_i)) {
                best = bv[t];
                best_i = bi[t];
            }
        }
        if (best_i < 0 || !(best > 0.0)) break;  // hvi.py:38-39: the caller draws the random fallback
        h_idx[b] = best_i;
        if (h_contrib) h_contrib[b] = best;
        const uint8_t one = 1;
        MCU(cudaMemcpy(dE.as<uint8_t>() + best_i, &one, 1, cudaMemcpyHostToDevice));
        for (int t = 0; t < m; ++t) front.push_back(h_Yc[best_i * m + t]);  // hvi.py:42
    }
    return AOED_OK;
}

}  // extern "C"
