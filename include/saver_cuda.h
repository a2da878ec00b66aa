/*
 * saver_cuda.h -- C ABI of libsaver_cuda, the B200 (sm_100a) implementation of SAVER's
 * per-gene expression-recovery hot path.
 *
 * The reference (mohuangx/SAVER, /root/reference) is pure R and has no FFI of its own; the
 * boundary it exposes for this path is its R function API (NAMESPACE:3-23).  Each entry
 * point below is what a `.Call` shim binds to replace the body of one of those functions
 * (INTEGRATION.md shows the shim).  Conventions:
 *   - plain C types only; no R, torch or CUDA types cross the boundary;
 *   - matrices are column-major fp64 like R's: a "panel" of B genes over n cells is n x B
 *     (one gene = one contiguous column); the design is n x p;
 *   - all buffers are caller-owned; the library keeps no host pointer past return;
 *   - every function returns 0 on success or a negative SAVER_ERR_* code;
 *     saver_cuda_last_error() gives the message.  Per-gene solver trouble is NOT an error:
 *     it is reported in per-gene status arrays, and the host mirrors the reference's
 *     tryCatch fallbacks (R/expr_predict.R:44-52, R/calc_posterior.R:41-49);
 *   - `device_ptrs` != 0 means every array argument of that call already lives on the
 *     handle's device (used by bench.py's HBM-resident timing and by the NCCL plumbing);
 *     an R host always passes 0;
 *   - there is no CPU fallback: without a usable CUDA device saver_cuda_create fails.
 */
#ifndef SAVER_CUDA_H
#define SAVER_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SAVER_API __attribute__((visibility("default")))
#else
#define SAVER_API
#endif

#define SAVER_ABI_VERSION 1

#define SAVER_OK 0
#define SAVER_ERR_CUDA (-1)     /* CUDA runtime failure (message has the CUDA error) */
#define SAVER_ERR_ARG (-2)      /* invalid argument */
#define SAVER_ERR_STATE (-3)    /* call order (e.g. design not set) */
#define SAVER_ERR_NOMEM (-4)    /* device or host allocation failed */

/* model codes of calc.loglik.* / calc.a|b|k */
#define SAVER_MODEL_A 0
#define SAVER_MODEL_B 1
#define SAVER_MODEL_K 2

/* per-gene info record written by saver_cuda_calc_post (doubles) */
#define SAVER_POST_INFO 8
/*  [0] chosen model: 0 a, 1 b, 2 k, 3 constant-mu branch (calc.b only), -1 all-zero gene
 *  [1] prior parameter of the chosen model (1/a, 1/b or k)
 *  [2..4] negative log-likelihood of a, b, k (NaN if not fitted)
 *  [5] bitmask: model m took the grid ("stochastic fallback") branch
 *  [6] number of log-likelihood evaluations
 *  [7] bitmask: model m raised an R error (dropped, as tryCatch -> c(0, Inf))            */

typedef struct saver_cuda_ctx* saver_cuda_handle;

/* ---- lifecycle ---- */
SAVER_API int saver_cuda_abi_version(void);
/* One handle per CUDA device (one process per GPU). */
SAVER_API int saver_cuda_create(int device, saver_cuda_handle* out);
/* Several devices behind one handle, for hosts that are ONE process (an R session): the in-process
 * analogue of the reference's ncores = PSOCK workers (R/saver.R:120-138).  devs[ndev]: CUDA device
 * ordinals.  On such a handle saver_cuda_set_design installs the design on every device (uploaded
 * once, then copied device to device) and saver_cuda_calc_estimate splits the block's genes into
 * contiguous chunks, one per device, on one host thread each (host buffers only); every other entry
 * point runs on devs[0].  Python hosts use one process per GPU instead (saver_b200/dist.py). */
SAVER_API int saver_cuda_create_multi(int ndev, const int* devs, saver_cuda_handle* out);
SAVER_API int saver_cuda_device_count(saver_cuda_handle h);
SAVER_API int saver_cuda_destroy(saver_cuda_handle h);
SAVER_API const char* saver_cuda_last_error(saver_cuda_handle h);
/* Block until all work queued by this handle has finished. */
SAVER_API int saver_cuda_synchronize(saver_cuda_handle h);
/* The cudaStream_t the handle launches on (as void*), for callers that time with events. */
SAVER_API int saver_cuda_stream(saver_cuda_handle h, void** stream);
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
SAVER_API long long saver_cuda_launch_count(saver_cuda_handle h);

/* ---- calc.loglik.a/b/k (R/calc_loglik.R:26,44,61): out = -loglik(param) ----
 * mu_len / sf_len are 1 (recycled scalar, R/calc_loglik.R:28-33) or n.  Host pointers. */
SAVER_API int saver_cuda_loglik(saver_cuda_handle h, int model, double param, const double* y,
                      const double* mu, int mu_len, const double* sf, int sf_len, int n,
                      double* out);

/* ---- calc.a / calc.b / calc.k (R/optimize_variance.R:24,61,101) ----
 * out2 = c(1/a | 1/b | k, nll).  flags: bit0 grid branch, bit1 uniroot used, bit2 the R
 * code would have raised (out2 = (0, Inf)).  grid200 (optional): the 100 grid points then
 * their 100 log-likelihoods, so an R host can draw sample() itself (optimize_variance.R:53);
 * the library uses the expectation sum(samp * p).  Host pointers. */
SAVER_API int saver_cuda_calc_var(saver_cuda_handle h, int model, const double* y, const double* mu,
                        int mu_len, const double* sf, int sf_len, int n, double* out2,
                        int* flags, double* grid200);

/* ---- calc.post (R/calc_posterior.R:26) for a panel of B genes ----
 * Y: n x B raw counts.  MU: n x B predictions, or B scalars when mu_is_scalar != 0
 * (R/calc_posterior.R:28-30).  sf: n size factors.  est (required), se (NULL when
 * estimates.only): n x B, ceiling(v * 1000 * scale_sf) / 1000.  est_raw / se_raw (optional):
 * the same before the ceiling.  info (optional): SAVER_POST_INFO x B.  grid (optional):
 * 600 x B, per model 100 grid points + 100 log-liks (only filled for grid-branch models). */
SAVER_API int saver_cuda_calc_post(saver_cuda_handle h, const double* Y, const double* MU,
                         int mu_is_scalar, const double* sf, int n, int B, double scale_sf,
                         double* est, double* se, double* est_raw, double* se_raw,
                         double* info, double* grid, int device_ptrs);

/* ---- the shared design: x.est = t(log((x[good,] + 1) / sf)) of R/saver.R:156-157 ----
 * xest: n x p column-major (n cells, p predictor genes).  Uploaded (or taken from the
 * device), standardised once and kept resident in the handle until replaced/destroyed. */
SAVER_API int saver_cuda_set_design(saver_cuda_handle h, const double* xest, int n, int p,
                                    int device_ptrs);

/* ---- calc.maxcor (R/calc_maxcor.R:26): out[g] = max_j |cor(x.est[, j], Y[, g])| ----
 * Y: n x B (size-normalised expression, R/calc_estimate.R:75-77).  self_col[g] = 0-based
 * column of x.est carrying the same gene name (zeroed pair), or -1; may be NULL. */
SAVER_API int saver_cuda_maxcor(saver_cuda_handle h, const double* Y, int B, const int* self_col,
                                double* out, int device_ptrs);

/* ---- expr.predict, CV variant (R/expr_predict.R:39-60) for a panel of B genes ----
 * = cv.glmnet(x.est[pred.cells, -self], y[pred.cells], family = "poisson", dfmax, nfolds),
 * predict(s = "lambda.min", type = "response") for ALL cells.
 * Y: n x B size-normalised expression.  foldid: n x B ints, 0 = cell is not a pred cell,
 * 1..nfolds = the fold cv.glmnet drew for that cell (the R host draws it: set.seed(seed);
 * sample(rep(seq(nfolds), length = N)), R/expr_predict.R:35-36 + cv.glmnet).  self_col as in
 * maxcor.  mu: n x B.  fit4: 4 x B = lambda.max, lambda.min, sd.cv, 0-based index of lambda.min.
 * status[B]: 0 ok; 1 sd(y) == 0 (mu = mean, R/expr_predict.R:37-38); 2 the fit failed and mu is
 * the mean over pred cells with lambda.max = lambda.min = sd.cv = 0 (R/expr_predict.R:44-52).
 * stats (optional) 4 x B ints: master path length, glmnet jerr, CD sweeps, gradient rounds.
 * cvout (optional) 3 x 128 x B: lambda, cvm, cvsd of cv.glmnet (NaN beyond the path). */
SAVER_API int saver_cuda_predict_cv(saver_cuda_handle h, const double* Y, int B, const int* self_col,
                                    const int* foldid, int nfolds, int dfmax, int nlambda,
                                    double thresh, double* mu, double* fit4, int* status,
                                    int* stats, double* cvout, int device_ptrs);

/* ---- expr.predict, fixed-lambda variant (R/expr_predict.R:61-82) ----
 * glmnet(lambda = c(exp(seq(log(lmax), log(lmin), by = -0.2)), lmin)), mu at s = lmin.
 * pred_mask: n ints (non-zero = pred cell) or NULL for all cells.  lambda_max / lambda_min:
 * HOST arrays of B values from est.lambda (R/utils.R:129-134) whatever device_ptrs says. */
SAVER_API int saver_cuda_predict_path(saver_cuda_handle h, const double* Y, int B,
                                      const int* self_col, const int* pred_mask,
                                      const double* lambda_max, const double* lambda_min,
                                      int dfmax, double thresh, double* mu, int* status,
                                      int* stats, int device_ptrs);

/* ---- calc.estimate (R/calc_estimate.R:65-161): the data-parallel loop body for a block ----
 * X: n x B raw counts of the block (one gene per column); sf: n normalised size factors.
 * self_col[B]: column of x.est with the gene's own name or -1 (R/calc_estimate.R:103);
 * is_pred[B]: the gene's name is in pred.gene.names (NULL = all);  pred_mask: n ints, non-zero =
 * pred cell (NULL = all);  foldid: n x B per-gene CV folds 1..nfolds (mode 1 only; the R host
 * draws them with set.seed(j); sample(rep(seq(nfolds), length = N)) exactly as cv.glmnet would).
 * mode: 0 = null.model (means), 1 = coefs NULL (cross-validated lasso), 2 = coefs given
 * (est.lambda + fixed-lambda path).  calc_maxcor / cutoff / coefs[2] as in the R signature.
 * est, se (NULL when estimates.only): n x B.  info6: six vectors of B doubles back to back =
 * maxcor, lambda.max, lambda.min, sd.cv, pred.time, var.time (seconds per gene, device time of
 * the block / B).  mu_out (optional): n x B predictions.  status_out (optional): per gene
 * 0 ok / 1 sd(y)==0 / 2 fit failed -> mean (the reference's tryCatch fallbacks). */
SAVER_API int saver_cuda_calc_estimate(saver_cuda_handle h, const double* X, int n, int B,
                                       const double* sf, double scale_sf, const int* self_col,
                                       const int* is_pred, const int* pred_mask, const int* foldid,
                                       int nfolds, int mode, int calc_maxcor, double cutoff,
                                       const double* coefs, int dfmax, int nlambda, double thresh,
                                       double* est, double* se, double* info6, double* mu_out,
                                       int* status_out, int device_ptrs);

/* ---- sample.saver (R/sample_saver.R:32-79) for a panel of B genes ----
 * est, se: n x B.  nb_mode = 0: rgamma(shape = est^2/se^2, rate = est/se^2) rounded to 3
 * decimals; 1 (efficiency.known): rnbinom(mu = est, size = est^2/se^2).  Counter-based generator:
 * the draw for (gene, cell) depends only on (seed, rep_idx, gene0 + gene, cell), so any blocking or
 * sharding of the genes gives the same dataset.  gene0 = index of the panel's first gene. */
SAVER_API int saver_cuda_sample(saver_cuda_handle h, const double* est, const double* se, int n,
                                int B, long long gene0, int nb_mode, unsigned long long seed,
                                int rep_idx, double* out, int device_ptrs);

/* ---- cor.genes / cor.cells (R/cor_adjust.R:26-66) ----
 * est, se: n x G panels (one gene per column).  by_cells = 0: gene-gene (M = G vectors), 1:
 * cell-cell (M = n).  out: M x ncols column block [col0, col0 + ncols) of
 * cor(.) * outer(adj, adj), adj = sqrt(var / (var + mean(se^2))); col0 = 0, ncols = M for the
 * whole matrix (ranks of a multi-GPU job take disjoint column blocks).  cor_in (optional): the
 * same column block of a user-supplied cor.mat, used instead of computing the correlation. */
SAVER_API int saver_cuda_cor_adjust(saver_cuda_handle h, const double* est, const double* se, int n,
                                    int G, int by_cells, const double* cor_in, int col0, int ncols,
                                    double* out, int device_ptrs);

/* ---- cor.genes with the genes spread over ranks (R/cor_adjust.R:32-42 on a gene-sharded saver object) ----
 * saver_cuda_cor_prep: est, se n x B (this rank's genes) -> Z (ldz x B, ldz = n rounded up to 16,
 * rows beyond n zeroed): centred unit-norm rows, so that Z^T Z = cor(t(est)); adj (B) =
 * sqrt(var / (var + mean(se^2))).  The ranks exchange Z / adj (one all-gather), then
 * saver_cuda_cor_block writes out[row0 .. row0+nrows, c] (column-major, leading dimension ldo, c
 * counted from col0) = clamp(Z_r . Z_c) * adj_r * adj_c for one rectangular block, from the
 * gathered panel (device pointers; Z allocated and zeroed up to round_up(M, 128) + 64 columns).
 * The matrix is symmetric: a rank needs only the blocks on or above its place in a cyclic order. */
SAVER_API int saver_cuda_cor_prep(saver_cuda_handle h, const double* est, const double* se, int n,
                                  int B, double* Z, int ldz, double* adj, int device_ptrs);
SAVER_API int saver_cuda_cor_block(saver_cuda_handle h, const double* Z, int ldz, const double* adj,
                                   int M, int row0, int nrows, int col0, int ncols, double* out,
                                   int ldo);

/* ---- sparse counts: R's dgCMatrix (genes x cells, compressed by cell: slots p, i, x) ----
 * Uploaded once and kept resident.  Values < 0.001 are zeroed on upload (clean.data,
 * R/utils.R:11-19).  Host pointers. */
SAVER_API int saver_cuda_set_counts_csc(saver_cuda_handle h, const int* indptr, const int* indices,
                                        const double* data, int G, int n);
/* colSums (n) and rowSums (G) of the resident matrix: calc.size.factor (R/utils.R:72), the
 * rowMeans >= 0.1 filter (R/saver.R:156), get.pred.genes (R/utils.R:100-115).  Host outputs. */
SAVER_API int saver_cuda_csc_sums(saver_cuda_handle h, double* col_sums, double* row_sums);
/* Dense gene-major panel of B genes (as.matrix(x[block, ]) of R/calc_estimate.R:69): out is n x B.
 * transform 0: raw counts; 1: log((x + 1) / sf), i.e. rows of x.est (R/saver.R:157).  gene_idx and
 * sf are host arrays; out lives on the device when out_on_device != 0 (feed it to
 * saver_cuda_set_design / saver_cuda_calc_estimate with device_ptrs = 1). */
SAVER_API int saver_cuda_csc_gather(saver_cuda_handle h, const int* gene_idx, int B, const double* sf,
                                    int transform, double* out, int out_on_device);

/* ---- dense counts: R's plain matrix, the hot path's a1-a3 (R/saver.R:105-157, R/utils.R:2-31,70-86) ----
 * x: genes x cells as G contiguous rows of n cells (= t(x) column-major, the panel layout), dtype 0 =
 * double, 1 = int32 (host pointer).  Uploaded once and kept resident until replaced or released;
 * values < 0.001 read as 0 (clean.data, R/utils.R:11-19).  The sums feed calc.size.factor
 * (R/utils.R:72), the rowMeans >= 0.1 filter (R/saver.R:156) and get.pred.genes; the gather with
 * transform = 1 is x.est = log((x[good, ] + 1) / sf) (R/saver.R:157) written straight into a device
 * panel for saver_cuda_set_design, so the design never exists on the host. */
SAVER_API int saver_cuda_set_counts_dense(saver_cuda_handle h, const void* x, int dtype, int G, int n);
SAVER_API int saver_cuda_dense_sums(saver_cuda_handle h, double* col_sums, double* row_sums);
SAVER_API int saver_cuda_dense_gather(saver_cuda_handle h, const int* gene_idx, int B, const double* sf,
                                      int transform, double* out, int out_on_device);
SAVER_API int saver_cuda_release_counts(saver_cuda_handle h);

/* Problem columns pushed through the gradient GEMM so far (roofline accounting in bench.py):
 * every column is one 2 * n * p flop contraction against the resident design. */
SAVER_API long long saver_cuda_gemm_columns(saver_cuda_handle h);

/* Cumulative accounting since the handle was created (CUDA-event device times, ms):
 * [0] lasso round kernels, [1] gradient GEMMs of the rounds, [2] lasso block setup (moments,
 * initial gradient), [3] variance/posterior kernel inside calc_estimate, [4] design columns the
 * round kernels streamed (n x 8 bytes each), [5] GEMM problem columns, [6] rounds, [7] launches. */
SAVER_API int saver_cuda_counters(saver_cuda_handle h, double* out8);
/* [0] sum of round-kernel CTA lifetimes (SM cycles), [1] the longest single CTA; [2..3] reserved */
SAVER_API int saver_cuda_counters_ext(saver_cuda_handle h, double* out4);
/* Summed SM cycles of the Gram round kernel per phase: [0] gradient/KKT/strong rule, [1] gradient
 * vectors, [2] Gram block (DMMA), [3] Gram-space sweeps, [4] residual materialisation, [5] strong-set
 * check, [6] re-linearisation, [7] number of IRLS steps, [8] Gram-space coordinate steps, [9] sweeps,
 * [10] strong-set columns checked, [11] checks, [12] sum of |A| over IRLS steps; out has 16 slots. */
SAVER_API int saver_cuda_phase_cycles(saver_cuda_handle h, double* out16);
/* Gram-space coordinate steps by active-set size (20 buckets of 32), counted with lasso_dbg bit 128. */
SAVER_API int saver_cuda_sweep_hist(saver_cuda_handle h, double* out20);

/* Solver options.  "lasso_gram" (default 1): 1 = Gram-space round kernel; 0 = the n-space kernel
 * that follows glmnet's fishnet1 sweep for sweep (slower; agrees with the CPU oracle to rounding on
 * most genes).  "strong_slack" (default 1 = sequential strong rule).  "lasso_dbg": bring-up bits. */
SAVER_API int saver_cuda_set_option(saver_cuda_handle h, const char* key, double value);

/* Diagnostics: the master path of gene g of the LAST block the solver ran -- per lambda (128
 * slots) intercept, deviance ratio, ever-active count, cumulative sweeps, lambda. */
SAVER_API int saver_cuda_debug_path(saver_cuda_handle h, int g, double* a0, double* dev, int* nin,
                                    int* nlp, double* lam);

#ifdef __cplusplus
}
#endif
#endif /* SAVER_CUDA_H */
