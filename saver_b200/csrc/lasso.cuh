// lasso.cuh -- data structures of the batched Poisson-lasso path solver (lasso.cu).
#pragma once
#include <cuda_runtime.h>

namespace saver {

constexpr int kNX = 640;     // capacity of the ever-active list (glmnet pmax = min(2*dfmax+20, nvars))
constexpr int kNLam = 128;   // capacity of a lambda sequence (glmnet nlambda = 100)
constexpr int kMaxFolds = 16;
constexpr int kCrumbSlots = 16384;
constexpr int kGramItems = 256;  // (tile, row segment) items of one row-split Gram build

// ST_HOLD: a fold fit that would get more than one lambda ahead of its gene's master fit waits (it
// keeps its place in the problem list and is looked at again next round).
enum ProbState : int { ST_INIT = 0, ST_SOLVE = 1, ST_WAIT_KKT = 2, ST_DONE = 3, ST_HOLD = 4 };

// One (gene, fit) problem: fit 0 = master fit on all pred cells, fit f > 0 = CV fold f held out.
struct LassoProb {
  int gene, fit;
  int state;
  int lidx;     // lambda index being solved
  int lmu;      // number of committed solutions
  int nin;      // ever-active count
  int nS;       // strong-set size (length of the problem's Sidx list)
  int jerr;     // glmnet code: 0 ok, -ilm maxit, -10000-ilm pmax overflow, >0 fatal
  int nlp;      // coordinate-descent sweeps so far
  int ngrad;    // full-gradient (X^T r over all p) rounds so far
  int mtrain;   // training rows
  int slot;     // column of R / G this problem used in the last round
  double al, al0;   // current / previous lambda
  double az;        // intercept
  double qv;        // observation weight 1 / mtrain
  double yb, dv0, dvr, shr;
  double v0;        // sum of working weights
  double sumr;      // sum of working residuals in R (standardisation shift of X^T r)
  double cur_dev;   // deviance ratio of the current iterate
  double cur_cv;    // held-out Poisson deviance (sum) of the current iterate
  double ncv;       // held-out rows
};

struct LassoParams {
  int B;            // genes in this block
  int NF;           // fits per gene: 1 (path variant) or nfolds + 1 (CV variant)
  int nfolds;
  int auto_lambda;  // 1: glmnet's automatic sequence (master) + folds on the master grid; 0: user
  int dfmax;        // glmnet dfmax (ne)
  int nlam;         // nlambda for the automatic sequence
  double thr;       // glmnet thresh
  int maxit;        // glmnet maxit
  double alf_wide;  // lambda ratio for flmin = 0.01 (nobs < nvars), computed on the host
  double alf_tall;  // lambda ratio for flmin = 1e-4
  int gram = 1;     // 1: Gram-space round kernel (default), 0: n-space kernel that mirrors fishnet1's iterate sequence
  double strong_slack = 1.0;  // screening margin below lambda in units of the lambda step (1 = strong rule)
  double inner_scale = 1.0;   // Gram kernel: the sweeps inside one IRLS step stop at thr * inner_scale * dev0
  double chord_tol = 1.0e4;   // Gram kernel: an IRLS step that moved a coefficient by more than chord_tol x
                              // the convergence threshold invalidates the kept curvature model
  int dbg = 0;      // bisection switches for bring-up (SAVER_LASSO_DBG), 0 in production
};

// Developer overrides of the solver's launch plan.  Read from the environment ONCE, when the handle
// is created (saver_cuda_create); nothing on the per-block path calls getenv.
struct LassoTuning {
  int problems = 0;        // SAVER_LASSO_PROBLEMS: problems in flight per block (0 = default)
  int dbg = -1;            // SAVER_LASSO_DBG: bring-up bits (-1 = use the handle option)
  int gram = -1;           // SAVER_LASSO_GRAM: force the round kernel variant
  int threads = 0;         // SAVER_LASSO_THREADS / SAVER_GRAM_THREADS: CTA size
  int gram_threads = 0;
  int gram_ctas = 0;       // SAVER_GRAM_CTAS: resident CTAs per SM to plan shared memory for
  int gram_staged = -1;    // SAVER_GRAM_STAGED: 0 = working weights / residual stay in global memory
  long gram_hs = -1;       // SAVER_GRAM_HS: cap (bytes) of the packed Gram block in shared memory
  int l2_persist = -1;     // SAVER_L2_PERSIST: 0 = no persisting-L2 window on the design
  double strong_slack = -1.0;  // SAVER_STRONG_SLACK
  double chord_tol = -1.0; // SAVER_CHORD_TOL
  int f32 = -1;            // SAVER_LASSO_F32: 0 = the curvature model reads the FP64 design too
  int order = -1;          // SAVER_LASSO_ORDER: 0 completion order, 1 fits of a gene adjacent, 2 + one cluster per gene
  void from_env();
};

// Device workspace of one block of problems (owned by the handle, grown on demand).
struct LassoWs {
  LassoTuning tune;
  size_t cap_bytes = 0;
  unsigned char* base = nullptr;
  int *h_counter = nullptr;  // pinned host mirror of the active-problem counter
  // fault breadcrumbs (lasso_dbg bit 4096): mapped pinned host memory that survives a device fault;
  // kCrumbSlots records of 8 ints, one per resident CTA slot, rewritten at every phase boundary
  int *h_crumb = nullptr, *d_crumb = nullptr;
  // carved pointers (valid for the current block)
  double *R, *W, *WR;        // ldx x P_pad panels: GEMM operand, working weights, working residual
  double *G, *XM, *XI;       // p_pad x P_pad tables: gradient, per-fit mean and 1/sd of Xs columns
  int *MM, *SIDX;            // p_pad x P: 0 none, -1 strong only, >0 active position / strong list
  int *IA;                   // kNX x P ever-active variables in order of entry
  double *AACT;              // kNX x P coefficients in active-list order (standardised scale)
  double *AS;                // kNX x P coefficients at the start of the current IRLS step (gram kernel)
  LassoProb* prob;           // P
  double *lam;               // kNLam x B
  int *nlam_g, *gstat;       // B
  int *msnap;                // 2 x B: the master fit's committed count and (when done) its final length,
                             // as of the START of the round -- what the fold fits steer by
  double *gmean;             // B: mean of y over the pred cells
  double *a0p, *devp;        // kNLam x B master path
  int *ninp;                 // kNLam x B
  int *nlpp;                 // kNLam x B sweeps so far at each commit (diagnostics)
  double *cap;               // kNX x kNLam x B master coefficients per lambda
  double *cvdev;             // kNLam x kMaxFolds x B held-out mean deviance
  int *list0, *list1;        // P each: active problem lists (ping-pong)
  int *counters;             // [2]
  unsigned long long* colcount;  // [0] design columns streamed, [1] CTA cycles, [2] max, [3] exits, [4] Gram FMAs
  int *INACT;                // p_pad x P: inactive members of the strong set (gram kernel)
  int *hflags;               // Gram scratch slot flags
  double *Hs;                // Gram scratch: kHSlots x kNX x kNX
  double *HP;                // partial Gram tiles of row-split builds: kHSlots x kGramItems x 256
  unsigned long long* phasecyc;  // [48] per-phase SM cycles of the gram kernel (bring-up / profiles)
  // accounting (host side, cumulative over the handle's life)
  double ms_round = 0, ms_gemm = 0, ms_setup = 0;
  unsigned long long cols_streamed = 0, cta_cycles = 0, max_cta_cycles = 0, gram_fma = 0;
  long long rounds = 0;
  unsigned long long phase_cycles[48] = {0};  // [16..35]: sweep steps by active-set size / 32
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  void release();
};

}  // namespace saver
