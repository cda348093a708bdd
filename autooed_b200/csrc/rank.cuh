// One-CTA non-dominated sort shared by the Pareto C ABI (moo.cu) and the device-resident NSGA-II (nsga2.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>

namespace aoed {

// Programmatic dependent launch (PDL): a kernel launched with `launch_pdl` may be scheduled while its predecessor in the stream
// is still running; it must call pdl_wait() before it touches anything the predecessor wrote (the wait returns once the
// predecessor grid has completed and its writes are visible).  pdl_trigger() in the predecessor allows that early scheduling.
// Both are no-ops in a kernel launched the ordinary way.  In the NSGA-II generation loop this hides the launch latency of the
// short kernels that follow each other (~2 us per boundary).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Whole non-dominated sort in ONE CTA for populations that fit in shared memory (the NSGA-II regime, n <= 4096):
// domination counts, then front peeling with the counts decremented by the members of each peeled front — no host
// round trip per front.  Same semantics as the multi-launch path below (rank = front number, pymoo NonDominatedSorting).
constexpr int RS_THREADS = 1024;

// slot = (*ctr)++ for the lanes with `pred`, one shared-memory atomic per warp instead of one per lane (a thousand
// same-address atomics serialise).  Must be reached by all 32 lanes; returns -1 for lanes without `pred`.
__device__ __forceinline__ int warp_slot(int* ctr, bool pred) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0) return -1;
    const int lane = threadIdx.x & 31, leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    return pred ? base + __popc(mask & ((1u << lane) - 1u)) : -1;
}
constexpr int RS_MAX_N = 4096;
constexpr int RS_PT = RS_MAX_N / RS_THREADS;  // points per thread of the one-CTA sort

// domination counts of all n points over many CTAs: a warp owns 32 consecutive points; the candidates for "dominates
// it" are cut into DC_SPLIT contiguous slices (grid.y), staged in shared memory — read straight from global memory every
// step of the scan is a fresh cache line and pays an L2 round trip — and scanned by the DC_WARPS warps of the CTA
// (every lane reads the same j: broadcast).  cnt[y * n + i] = partial count of slice y; the sort kernel adds the slices.
constexpr int DC_WARPS = 8;
constexpr int DC_SPLIT = 4;
inline int dom_count_chunk(int n) { return (n + DC_SPLIT - 1) / DC_SPLIT; }

template <int M>
__global__ void __launch_bounds__(32 * DC_WARPS) dom_count_kernel(const double* __restrict__ YT, int n, int chunk, int* __restrict__ cnt) {
    extern __shared__ double dc_smem[];  // [M][chunk]
    __shared__ int part[32];
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, i = blockIdx.x * 32 + lane;
    const int j0 = blockIdx.y * chunk, len = min(chunk, n - j0);
    for (int t = threadIdx.x; t < M * chunk; t += 32 * DC_WARPS) {
        const int k = t / chunk, j = t - k * chunk;
        if (j < len) dc_smem[t] = YT[int64_t(k) * n + j0 + j];
    }
    if (w == 0) part[lane] = 0;
    __syncthreads();
    if (i < n) {
        double yi[M];
#pragma unroll
        for (int k = 0; k < M; ++k) yi[k] = YT[int64_t(k) * n + i];
        int c = 0;
#pragma unroll 4
        for (int j = w; j < len; j += DC_WARPS) {
            bool le = true, lt = false;
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const double v = dc_smem[k * chunk + j];
                le = le && (v <= yi[k]);
                lt = lt || (v < yi[k]);
            }
            c += (le && lt) ? 1 : 0;
        }
        atomicAdd(&part[lane], c);
    }
    __syncthreads();
    if (w == 0 && i < n) cnt[int64_t(blockIdx.y) * n + i] = part[lane];
}

// `cnt_in` (optional): domination counts from dom_count_kernel — worth the extra launch for n above ~1000, where the
// single CTA would spend most of its time on the n^2 counting.  `need`: stop once that many points are ranked (NSGA-II
// survival only needs the fronts that fill the next population, pymoo `n_stop_if_ranked`); the rest get the next rank
// number.  need >= n gives the full sort.  `last_rank_out` (optional) receives the last rank that is a real front.
template <int M>
__global__ void __launch_bounds__(RS_THREADS) rank_small_kernel(const double* __restrict__ YT, int n, int32_t* __restrict__ rank_out,
                                                                const int* __restrict__ cnt_in, int need,
                                                                int* __restrict__ last_rank_out) {
    extern __shared__ double rs_smem[];
    double* Y = rs_smem;                                    // [M][n]
    int* list0 = reinterpret_cast<int*>(Y + size_t(M) * n); // [n] compacted fronts, alternating: the one being peeled ...
    int* list1 = list0 + n;                                 // [n] ... and the one that forms while it is peeled
    __shared__ int n_list[3];  // member counts, rotating: read in round r, filled in round r (for r + 1), cleared for r + 2
    const int tid = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    for (int i = tid; i < M * n; i += RS_THREADS) Y[i] = YT[i];
    if (tid < 3) n_list[tid] = 0;
    __syncthreads();
    // A thread owns the points tid, tid + 1024, ... (at most RS_PT): their coordinates, remaining-dominator counts and ranks
    // stay in registers for the whole sort; shared memory holds the coordinate table and the front lists only.
    // (1024 threads leave 64 registers each: the coordinates are cached for up to four objectives, read from shared memory beyond)
    constexpr bool CACHE = M <= 4;
    double yq[CACHE ? RS_PT : 1][M];
    int cnt[RS_PT], rk[RS_PT];
#pragma unroll
    for (int u = 0; u < RS_PT; ++u) {  // domination counts; the points nobody dominates are front 0
        const int i = u * RS_THREADS + tid;
        int c = 1;
        if (i < n) {
            if (CACHE) {
#pragma unroll
                for (int k = 0; k < M; ++k) yq[u][k] = Y[k * n + i];
            }
            c = 0;
            if (cnt_in) {
#pragma unroll
                for (int y = 0; y < DC_SPLIT; ++y) c += cnt_in[y * n + i];
            } else {
                for (int j = 0; j < n; ++j) {
                    bool le = true, lt = false;
#pragma unroll
                    for (int k = 0; k < M; ++k) {
                        const double v = Y[k * n + j], mine = CACHE ? yq[u][k] : Y[k * n + i];
                        le = le && (v <= mine);
                        lt = lt || (v < mine);
                    }
                    c += (le && lt) ? 1 : 0;
                }
            }
        }
        cnt[u] = c;
        rk[u] = c == 0 ? 0 : -1;
        if (u * RS_THREADS < n) {  // uniform over the CTA
            const int slot = warp_slot(&n_list[0], i < n && c == 0);
            if (i < n && c == 0) list0[slot] = i;
        }
    }
    // Peeling.  ONE barrier per front: while front r is subtracted from the counts of the points it dominates, a point whose
    // count reaches 0 IS a member of front r + 1 and appends itself to the other list — no separate scan for the next front.
    // (Early NSGA-II generations have hundreds of fronts of a few members each: the cost of a sort is the cost per front.)
    int left = n, last_full = 0;  // uniform over the CTA: every thread follows the same counts
    for (int r = 0;; ++r) {
        __syncthreads();
        const int nf = n_list[r % 3];
        if (nf == 0) break;  // with points left: only for non-finite inputs (a finite set always has a minimal element)
        if (n - left + nf >= need && nf < left) {  // enough ranked: everything left shares the next rank number
#pragma unroll
            for (int u = 0; u < RS_PT; ++u)
                if (rk[u] < 0) rk[u] = r + 1;
            last_full = r;
            break;
        }
        last_full = r;
        left -= nf;
        if (left == 0) break;
        if (tid == 0) n_list[(r + 2) % 3] = 0;  // last read a round ago (every thread has passed this round's barrier since)
        const int* front = (r & 1) ? list1 : list0;
        int* next = (r & 1) ? list0 : list1;
#pragma unroll
        for (int u = 0; u < RS_PT; ++u) {
            if (u * RS_THREADS >= n) break;  // uniform
            const int q = u * RS_THREADS + tid;
            bool joins = false;
            if (q < n && rk[u] < 0) {
                int dec = 0;
                for (int t = 0; t < nf; ++t) {
                    const int pidx = front[t];
                    bool le = true, lt = false;
#pragma unroll
                    for (int k = 0; k < M; ++k) {
                        const double v = Y[k * n + pidx], mine = CACHE ? yq[u][k] : Y[k * n + q];
                        le = le && (v <= mine);
                        lt = lt || (v < mine);
                    }
                    dec += (le && lt) ? 1 : 0;
                }
                cnt[u] -= dec;
                if (cnt[u] == 0) {
                    rk[u] = r + 1;
                    joins = true;
                }
            }
            const int slot = warp_slot(&n_list[(r + 1) % 3], joins);
            if (joins) next[slot] = q;
        }
    }
#pragma unroll
    for (int u = 0; u < RS_PT; ++u) {
        const int i = u * RS_THREADS + tid;
        if (i < n) rank_out[i] = rk[u];
    }
    if (last_rank_out && tid == 0) *last_rank_out = last_full;  // highest rank whose front was peeled completely
}


// device ints the rank launchers need as scratch for n points (the DC_SPLIT partial counts)
inline size_t rank_scratch_ints(int n) { return size_t(DC_SPLIT) * n; }

inline size_t rank_small_smem(int m, int n) { return size_t(m) * n * sizeof(double) + 2 * size_t(n) * sizeof(int); }
constexpr int RS_SPLIT_N = 1024;  // above this the domination counts come from dom_count_kernel

// launches the sort of n <= RS_MAX_N points on stream s; `cnt_scratch` ([rank_scratch_ints(n)] ints on the device) enables the
// split
// counting over many CTAs for large n (pass nullptr to stay with the single kernel)
constexpr size_t RS_SMEM_LIMIT = 200 * 1024;  // callers check rank_small_smem(m, n) against this

template <int M>
cudaError_t launch_rank_small_m(const double* YT, int n, int32_t* rank, int* cnt_scratch, int need, cudaStream_t s,
                                int* last_rank_out = nullptr, bool pdl = false) {
    const size_t sm = rank_small_smem(M, n);
    const int chunk = dom_count_chunk(n);
    const size_t dsm = size_t(M) * chunk * sizeof(double);
    static thread_local int opted_in_device = -1;  // the shared-memory opt-in is per device and costs a driver call
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (opted_in_device != dev) {
        e = cudaFuncSetAttribute(rank_small_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(RS_SMEM_LIMIT));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(dom_count_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(size_t(M) * dom_count_chunk(RS_MAX_N) * sizeof(double)));
        if (e != cudaSuccess) return e;
        opted_in_device = dev;
    }
    const int* cnt_in = nullptr;
    if (cnt_scratch && n > RS_SPLIT_N) {
        e = launch_pdl(pdl, dom_count_kernel<M>, dim3((n + 31) / 32, DC_SPLIT), dim3(32 * DC_WARPS), dsm, s, YT, n, chunk, cnt_scratch);
        if (e != cudaSuccess) return e;
        cnt_in = cnt_scratch;
    }
    e = launch_pdl(pdl, rank_small_kernel<M>, dim3(1), dim3(RS_THREADS), sm, s, YT, n, rank, cnt_in, need, last_rank_out);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace aoed
