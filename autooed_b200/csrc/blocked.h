// Host interface of the batched blocked Cholesky / triangular inverse / K^-1 used by the hyper-parameter fit for
// training sets that do not fit one CTA's shared memory (n_train > ~230) and by aoed_gp_set_theta at every size.
// Implementation: blocked.cu (tgemm_kernel, potrf_diag_kernel).  All matrices are row-major [batch][Npad][Npad] with
// Npad a multiple of 128 and IDENTITY padding (K_pad = diag(K, I)), which keeps every tile of the algorithm regular.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace aoed {

struct TileDesc;

struct BlockedSeg {
    int off = 0, n = 0;
};

// Tile table of one matrix size (T = Npad / 128 tile rows), resident on the device, plus the launch segments into it.
struct BlockedPlan {
    int T = 0;
    TileDesc* dTiles = nullptr;
    std::vector<BlockedSeg> upd, pan;    // Cholesky step k: left-looking update of block column k, panel scaling
    std::vector<BlockedSeg> invA, invB;  // triangular inverse, recursive doubling level l: P = L21 X11, X21 = -X22 P
    BlockedSeg lauum;                    // K^-1 = L^-T L^-1, lower tiles
    int n_counters = 0;                  // ticket counters one full sequence consumes
};

// One batch of matrices with the TMA views the tile GEMMs need.
struct BlockedMats {
    int Npad = 0, cap = 0;
    double *K = nullptr;   // in: K (symmetric, identity padded); out: L in the lower triangle (upper tiles keep K)
    double *X = nullptr;   // out: L^-1 (lower tiles; tiles above the diagonal are never written)
    double *XT = nullptr;  // out: (L^-1)^T, formed by the caller's transpose between the two calls below
    double *P = nullptr;   // scratch of the inverse, then K^-1 (lower tiles)
    CUtensorMap K_a, X_a, X_b, XT_a, P_b;  // *_a: box 128 rows x 16 columns, *_b: box 16 x 16
};

int blocked_plan_build(BlockedPlan& plan, int T);
void blocked_plan_free(BlockedPlan& plan);
int blocked_mats_alloc(BlockedMats& mats, int Npad, int cap);
void blocked_mats_free(BlockedMats& mats);

// L = chol(K) in place and X = L^-1 for `batch` problems.  d_info[b] = 0 on entry; != 0 on exit marks a matrix that is
// not positive definite (1 + index of the failing pivot).  d_counters: plan.n_counters ints (zeroed here).
int blocked_chol_inverse(const BlockedPlan& plan, const BlockedMats& mats, int batch, int* d_info, int* d_counters,
                         int sms, cudaStream_t s, int64_t* n_launches);
// P = XT * X restricted to the lower tiles (K^-1); needs XT = X^T.
int blocked_kinv(const BlockedPlan& plan, const BlockedMats& mats, int batch, int* d_counters, int sms, cudaStream_t s,
                 int64_t* n_launches);

// defined in gp.cu: row-major FP64 matrix [rows][ld] viewed through boxes of box_rows x 16 columns, 128-byte swizzle
int make_tensor_map_2d(CUtensorMap* map, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

}  // namespace aoed

extern "C" int aoed_detail_fail(int code, const char* msg);
