/*
 * aoed_b200.h — C ABI of the B200-native GP-surrogate hot path of AutoOED.
 *
 * One shared library (libaoed_b200.so, built from autooed_b200/csrc/ for sm_100a) exports exactly
 * these entry points.  Plain pointers and sizes only: no torch / numpy types cross this boundary.
 * Every function returns AOED_OK (0) or a negative AOED_ERR_* code; aoed_last_error() gives the text.
 *
 * Conventions
 *   - all floating point is IEEE float64, all matrices are C-contiguous row-major (rows, features),
 *     exactly the numpy arrays the reference passes around (SURVEY.md §8b "Data conventions");
 *   - "normalised" X means clip((X - xl)/(xu - xl), 0, 1) and normalised Y means (Y - mean)/std, i.e.
 *     what autooed/utils/normalization.py:31-32,103-109 produces before `GaussianProcess._fit/_evaluate`;
 *   - theta is the sklearn log-parameter vector [log c1, log l_1..l_d, log c2] (p = d + 2) of the kernel
 *     c1 * Matern_nu(l) + c2 built at autooed/mobo/surrogate_model/gp.py:126-132; nu in {1,3,5,-1};
 *   - pointers named h_* are HOST pointers, d_* are DEVICE pointers on the handle's device.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the AutoOED tree).
 */
#ifndef AOED_B200_H
#define AOED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AOED_ABI_VERSION 1

/* return codes */
#define AOED_OK 0
#define AOED_ERR_INVALID (-1)   /* bad argument / unfitted model (the reference's AssertionError) */
#define AOED_ERR_CUDA (-2)      /* CUDA runtime failure (message in aoed_last_error) */
#define AOED_ERR_NOT_PD (-3)    /* Cholesky of K failed at the final theta (numpy.linalg.LinAlgError in sklearn _gpr.py:352-361) */
#define AOED_ERR_UNSUPPORTED (-4)
#define AOED_ERR_NO_DEVICE (-5) /* no CUDA device: there is NO CPU fallback */

/* evaluate flags (autooed/mobo/surrogate_model/base.py:68 `evaluate(X, std, gradient, hessian)`) */
#define AOED_EVAL_STD 1u
#define AOED_EVAL_GRAD 2u
#define AOED_EVAL_HESS 4u
#define AOED_EVAL_EXACT 8u /* exact derivatives instead of the reference's literal broadcasting */

/* acquisition kinds (autooed/mobo/factory.py:25-38) */
#define AOED_ACQ_NONE 0
#define AOED_ACQ_IDENTITY 1 /* acquisition/identity.py:15-18 */
#define AOED_ACQ_UCB 2      /* acquisition/ucb.py:19-53, param[0] = n_sample */
#define AOED_ACQ_EI 3       /* acquisition/ei.py:20-70,  param[0..m) = y_min */
#define AOED_ACQ_PI 4       /* acquisition/pi.py:20-64,  param[0..m) = y_min */

typedef struct aoed_gp aoed_gp_t; /* opaque: one multi-objective GP surrogate resident on one GPU */

/* ---- library ------------------------------------------------------------------------------ */
int aoed_abi_version(void);
const char* aoed_last_error(void);
int aoed_device_count(void);
/* pinned host memory so that result arrays can be copied back asynchronously */
int aoed_host_alloc(void** out, uint64_t bytes);
int aoed_host_free(void* p);

/* ---- model lifecycle ------------------------------------------------------------------------
 * replaces GaussianProcess.__init__ (gp.py:108-135): n_obj independent GPs, kernel selected by nu.
 * 1 <= n_var <= 128 (AOED_ERR_UNSUPPORTED beyond). */
int aoed_gp_create(aoed_gp_t** out, int device, int n_var, int n_obj, int nu);
int aoed_gp_destroy(aoed_gp_t* gp);

/* replaces the data hand-off of GaussianProcess._fit (gp.py:137-139) after SurrogateModel.fit normalised
 * the data (base.py:31-52): Xn (n_train x n_var), Yn (n_train x n_obj), plus the StandardScaler statistics
 * used by SurrogateModel.evaluate to un-normalise (base.py:106-109). */
int aoed_gp_set_train(aoed_gp_t* gp, int n_train, const double* h_Xn, const double* h_Yn,
                      const double* h_y_mean, const double* h_y_scale);

/* replaces sklearn GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True)
 * (_gpr.py:541-655, reached from gp.py:139) for a BATCH of `n_prob` independent problems at once:
 * problem b uses objective h_obj[b] and hyper-parameters h_theta[b*p .. b*p+p).
 * Outputs h_lml[b] (-inf when K is not PD, with zero gradient, as _gpr.py:589-593) and h_grad[b*p..]. */
int aoed_gp_lml_batch(aoed_gp_t* gp, int n_prob, const int32_t* h_obj, const double* h_theta,
                      double* h_lml, double* h_grad /* may be NULL */);

/* replaces the tail of GaussianProcessRegressor.fit (_gpr.py:349-367): fixes theta (n_obj x p), factorises
 * K = k(X,X) + 1e-10 I for all objectives in ONE batch of the library's own blocked Cholesky + triangular inverse
 * (FP64 DMMA tile GEMMs, csrc/blocked.cu — no cuSOLVER / cuBLAS), computes alpha_ and caches L^-1 for the posterior.
 * AOED_ERR_NOT_PD on failure. */
int aoed_gp_set_theta(aoed_gp_t* gp, const double* h_theta);

/* sklearn attributes other AutoOED modules read (`gp.L_`, `gp.alpha_`, `log_marginal_likelihood_value_`):
 * h_L (n_train x n_train lower, row-major), h_alpha (n_train); either may be NULL. */
int aoed_gp_get_state(aoed_gp_t* gp, int obj, double* h_L, double* h_alpha, double* h_lml);

/* ---- posterior + acquisition ------------------------------------------------------------------
 * replaces GaussianProcess._evaluate (gp.py:141-253) + the un-normalisation of SurrogateModel.evaluate
 * (base.py:106-109) + `<Acquisition>._evaluate` (identity.py / ucb.py / ei.py / pi.py) in ONE call.
 *   h_Xn      : n_rows x n_var normalised candidates
 *   flags     : AOED_EVAL_* (batched semantics = stack of per-row reference results)
 *   acq_kind  : AOED_ACQ_*; h_acq_param as documented above (NULL for NONE / IDENTITY)
 * Outputs (any may be NULL = not wanted; shapes as base.py:87-92):
 *   F, S  : n_rows x n_obj          dF, dS : n_rows x n_obj x n_var      hF, hS : n_rows x n_obj x n_var x n_var
 *   A, dA, hA : acquisition value / gradient / hessian with the same shapes as F, dF, hF
 * F, S, dF, dS are in raw-Y units; hF, hS stay in normalised-Y units unless AOED_EVAL_EXACT.
 * Host buffers; host<->device copies happen inside, chunked and overlapped with compute. */
int aoed_gp_eval(aoed_gp_t* gp, int64_t n_rows, const double* h_Xn, uint32_t flags, int acq_kind,
                 const double* h_acq_param, double* h_F, double* h_S, double* h_dF, double* h_dS,
                 double* h_hF, double* h_hS, double* h_A, double* h_dA, double* h_hA);

/* same computation with DEVICE-resident input and outputs (no PCIe traffic), enqueued on `stream`
 * (a cudaStream_t; NULL = default stream).  Used by bench.py's `value` leg and by multi-GPU sharding.
 * With AOED_EVAL_HESS the first-order quantities are always computed internally (the variance Hessian and the
 * EI / PI Hessians need them, gp.py:236-241), whatever AOED_EVAL_GRAD says. */
int aoed_gp_eval_device(aoed_gp_t* gp, int64_t n_rows, const double* d_Xn, uint32_t flags, int acq_kind,
                        const double* h_acq_param, double* d_F, double* d_S, double* d_dF, double* d_dS,
                        double* d_hF, double* d_hS, double* d_A, double* d_dA, double* d_hA, void* stream);

/* counters for bench.py: number of kernels this handle launched, and the accumulated device time (ms) of the
 * dominant kernel family (triangular FP64 GEMM) measured with CUDA events when profiling is switched on. */
int aoed_gp_set_profiling(aoed_gp_t* gp, int on);
int aoed_gp_get_counters(aoed_gp_t* gp, int64_t* n_launches, double* trmm_ms, int64_t* trmm_launches,
                         double* trmm_flops);
int aoed_gp_reset_counters(aoed_gp_t* gp);
/* how many log-marginal-likelihood problems each device path served: the one-CTA-per-problem shared-memory kernel
 * (n_train <= 232) and the batched blocked factorisation (csrc/blocked.cu: every launch covers all problems of the call). */
int aoed_gp_get_fit_counters(aoed_gp_t* gp, int64_t* n_small, int64_t* n_blocked);

/* ---- NSGA-II on the device -----------------------------------------------------------
 * replaces the pymoo NSGA2 + minimize(('n_gen', G)) loop of autooed/mobo/solver/nsga2.py:17-30 for unconstrained problems:
 * the generation loop (variation, duplicate elimination, acquisition evaluation through the fitted model `gp`,
 * non-dominated sorting, crowding, survival) runs on the GPU without host synchronisation.
 *   acq_kind / h_acq_param as for aoed_gp_eval (IDENTITY .. PI); h_X0 (n_init x n_var): initial sampling in CONTINUOUS
 *   space, taken as is (generation 1 = its evaluation); h_xl, h_xu (n_var): bounds; pymoo 0.4.1 operator defaults.
 * Outputs: h_X (pop_size x n_var), h_F (pop_size x n_obj) — the final population and its acquisition values (only the first
 * *h_n_out rows are valid: min(n_init, pop_size) when n_gen == 1, else pop_size); *h_n_eval = acquisition evaluations
 * that were not discarded as duplicates.  n_init + pop_size <= 4096. */
int aoed_nsga2_run(aoed_gp_t* gp, int device, int n_var, int n_obj, int acq_kind, const double* h_acq_param, int64_t n_init,
                   const double* h_X0, const double* h_xl, const double* h_xu, int pop_size, int n_gen, uint64_t seed,
                   double* h_X, double* h_F, int64_t* h_n_out, int64_t* h_n_eval);

/* the same loop on a Thompson-sampling posterior sample (arguments as aoed_rff_eval): lets TSEMO, the reference's default
 * algorithm ('gp' + 'ts' + 'nsga2' + 'hvi'), run its solver step on the device as well. */
int aoed_nsga2_run_rff(int device, int n_var, int n_obj, int n_feat, const double* h_W, const double* h_b, const double* h_theta,
                       const double* h_factor, const double* h_y_mean, const double* h_y_scale, int64_t n_init,
                       const double* h_X0, const double* h_xl, const double* h_xu, int pop_size, int n_gen, uint64_t seed,
                       double* h_X, double* h_F, int64_t* h_n_out, int64_t* h_n_eval);

/* one rank-and-crowding survival step of that loop on its own (pymoo 0.4.1 RankAndCrowdingSurvival, which the reference's
 * NSGA2 solver runs every generation — autooed/mobo/solver/nsga2.py:17-30): from the objective rows h_F (n x n_obj) keeps
 * the `keep` best.  Outputs: h_rank (n) front number, fronts are only peeled until `keep` individuals are ranked and every
 * individual beyond gets (last peeled front + 1); h_crowd (n) crowding distance inside the own front (0 for the unpeeled
 * rest); h_sel (keep) indices of the survivors ordered by (rank, crowding distance descending, index).  n <= 4096. */
int aoed_nsga2_survival(int device, int64_t n, int n_obj, const double* h_F, int64_t keep, int32_t* h_rank, double* h_crowd,
                        int32_t* h_sel);

/* ---- Thompson-sampling acquisition -----------------------------------------------
 * replaces ThompsonSampling._evaluate (autooed/mobo/acquisition/ts.py:65-87) incl. its un-normalisation: value, gradient
 * and Hessian of the random-Fourier-feature posterior sample  f_o(x) = factor_o * sum_j theta_oj cos(W_oj . x + b_oj).
 *   h_W (n_obj x n_feat x n_var), h_b, h_theta (n_obj x n_feat), h_factor = sqrt(2 sf2 / n_feat), h_y_mean, h_y_scale (n_obj)
 *   h_Xn (n_rows x n_var) normalised candidates; flags AOED_EVAL_GRAD / AOED_EVAL_HESS
 * Outputs (may be NULL): F = f * y_scale + y_mean (n_rows x n_obj); dF = df/dx * y_scale (n_rows x n_obj x n_var);
 * hF = d2f/dx2 in normalised-Y units, as the reference leaves it (n_rows x n_obj x n_var x n_var). */
int aoed_rff_eval(int device, int64_t n_rows, int n_var, int n_obj, int n_feat, const double* h_W, const double* h_b,
                  const double* h_theta, const double* h_factor, const double* h_y_mean, const double* h_y_scale,
                  const double* h_Xn, uint32_t flags, double* h_F, double* h_dF, double* h_hF);

/* ---- multi-objective selection ------------------------------------------------------------------
 * replaces find_pareto_front / check_pareto (autooed/utils/pareto.py:27-62): non-dominated mask of Y (n x m),
 * minimisation; i is kept iff no j has all(Y_j <= Y_i) and any(Y_j < Y_i). */
int aoed_pareto_mask(int device, int64_t n, int m, const double* h_Y, uint8_t* h_mask);
/* the same test for the row shard [row_lo, row_hi) against ALL n points (multi-GPU: rows sharded, columns replicated,
 * SURVEY.md §8e); h_mask has row_hi - row_lo entries. */
int aoed_pareto_mask_rows(int device, int64_t n, int m, const double* h_Y, int64_t row_lo, int64_t row_hi,
                          uint8_t* h_mask);
/* non-dominated sorting rank (0 = first front), pymoo NonDominatedSorting semantics. */
int aoed_pareto_rank(int device, int64_t n, int m, const double* h_Y, int32_t* h_rank);

/* replaces HypervolumeImprovement._select (autooed/mobo/selection/hvi.py:16-47): greedy batch of `batch`
 * candidate indices maximising the exclusive hypervolume w.r.t. the running front (initially h_front,
 * n_front x m), reference point h_ref (m).  Strict '>' so the lowest index wins ties (hvi.py:35-37);
 * when no candidate has a positive contribution the entry is -1 and the caller draws the random fallback
 * (hvi.py:38-39) with the host RNG, then calls again with the updated front (see selection/hvi.py in the
 * Python package).  h_contrib (batch) receives the winning contributions (may be NULL).
 * All rounds run on the device without host synchronisation (candidates, front and taken mask resident, one kernel per
 * round); a round whose two best contributions differ by <= 1e-12 (HV(front) + best) — or in which nothing is positive — is
 * re-decided on the host in the reference's own form HV(front + y) - HV(front) (hvi.py:28-37), pymoo's non-dominated
 * pre-filter included.  n_obj <= 8, n_obj * (n_front + batch) <= 25600. */
int aoed_hvi_select(int device, int64_t n_cand, int m, const double* h_Yc, int64_t n_front,
                    const double* h_front, const double* h_ref, int batch, const uint8_t* h_excluded,
                    int64_t* h_idx, double* h_contrib);
/* exclusive hypervolume contribution of every candidate w.r.t. a fixed front (one greedy round). */
int aoed_hv_contributions(int device, int64_t n_cand, int m, const double* h_Yc, int64_t n_front,
                          const double* h_front, const double* h_ref, double* h_contrib);

/* ---- sparse approximation of the Pareto front (ParetoDiscovery) -----------------------------------
 * replaces the third-party `pygco.cut_from_graph(edges, unary_cost, pairwise_cost, n_iter)` call of
 * autooed/mobo/solver/pareto_discovery/buffer.py:249 (alpha-expansion graph cut, gco-v3.0 behind a Cython wrapper):
 *   minimises  E(l) = sum_i unary[i][l_i] + sum_{(i,j) in edges} pairwise[l_i][l_j]
 * over labelings of n_node nodes with n_label labels, starting from label 0 everywhere, by cycles of expansion moves
 * (alpha = 0 .. n_label-1 in order, each move an exact minimum s-t cut, the minimum cut with the smallest sink side so
 * that ties never move a node); stops after a cycle without decrease or after n_iter cycles (n_iter < 0: until convergence).
 *   edges (n_edge x 2, node indices), unary (n_node x n_label), pairwise (n_label x n_label): int32 as pygco requires;
 *   labels (n_node) out; energy (may be NULL) = E of the returned labeling.
 * HOST routine (the graphs have one node per non-empty buffer cell): no device is touched.
 * AOED_ERR_UNSUPPORTED when a pair term is not submodular for a move (V(a,b) + V(c,c) > V(a,c) + V(c,b)). */
int aoed_graph_cut(int n_node, int64_t n_edge, const int32_t* edges, int n_label, const int32_t* unary,
                   const int32_t* pairwise, int n_iter, int32_t* labels, int64_t* energy);

#ifdef __cplusplus
}
#endif
#endif /* AOED_B200_H */
