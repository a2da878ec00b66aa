// lasso.cu -- batched cross-validated Poisson lasso paths on one shared design.
//
// Replaces (reference file:line):
//   R/expr_predict.R:39-60   expr.predict, CV variant   = glmnet::cv.glmnet(family="poisson",
//                            dfmax=300, nfolds=5), s = "lambda.min"
//   R/expr_predict.R:61-82   expr.predict, fixed-lambda variant = glmnet::glmnet(lambda=seq)
// glmnet itself is not in the reference tree; its dense Poisson path (Fortran fishnet1) and the
// cv.glmnet glue are restated in oracle/saver_oracle.c, which is what this
// file is checked against.
//
// Formulation.  Every (gene, fit) pair -- fit 0 = master fit on all pred cells, fit f = fold f
// held out -- is one "problem" with the SAME resident design Xs (cells x predictors, standardised
// once over all cells).  A fit's own standardisation (glmnet standardises per fit, with that
// fit's observation weights) is an affine map of Xs columns: x~ = (Xs - xm) * xi with per-fit
// moments xm, xi that come out of two GEMMs against the 0/1 fold-weight panel.  Held-out rows
// carry weight 0, so X is never row-compacted.
//
// Work split per lambda step of a problem:
//   * everything that touches only the strong set (IRLS re-linearisation, cyclic coordinate
//     descent, intercept) runs inside ONE CTA per problem with the working weights / residual
//     staged in shared memory (lasso_round_kernel);
//   * the full gradient X^T r over all p predictors -- glmnet's KKT check + strong rule, the
//     O(n p) part -- is batched over all live problems into one FP64 DMMA GEMM (gemm64.cu).
// A "round" = one GEMM + one round kernel; a problem advances one lambda (or repairs one KKT
// violation) per round.  Finished problems leave the panel, so the GEMM shrinks with them.
#include "lasso.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "saver_kernels.h"

namespace saver {

void LassoTuning::from_env() {
  auto geti = [](const char* k, int& v) { if (const char* e = getenv(k)) v = atoi(e); };
  geti("SAVER_LASSO_PROBLEMS", problems);
  geti("SAVER_LASSO_DBG", dbg);
  geti("SAVER_LASSO_GRAM", gram);
  geti("SAVER_LASSO_THREADS", threads);
  geti("SAVER_GRAM_THREADS", gram_threads);
  geti("SAVER_GRAM_CTAS", gram_ctas);
  geti("SAVER_GRAM_STAGED", gram_staged);
  geti("SAVER_L2_PERSIST", l2_persist);
  geti("SAVER_LASSO_ORDER", order);
  geti("SAVER_LASSO_F32", f32);
  if (const char* e = getenv("SAVER_GRAM_HS")) gram_hs = atol(e);
  if (const char* e = getenv("SAVER_STRONG_SLACK")) strong_slack = atof(e);
  if (const char* e = getenv("SAVER_CHORD_TOL")) chord_tol = atof(e);
}

void LassoWs::release() {
  if (base) cudaFree(base);
  if (h_counter) cudaFreeHost(h_counter);
  if (h_crumb) cudaFreeHost(h_crumb);
  h_crumb = d_crumb = nullptr;
  for (auto& e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
  base = nullptr;
  h_counter = nullptr;
  cap_bytes = 0;
}

namespace {

constexpr double kBig = 9.9e35, kSml = 1.0e-4 /* fdev 1e-5 x 10 */, kDevMax = 0.999;
constexpr int kMnLam = 5;
constexpr double kVarTiny = 1e-12;  // per-fit variance of a unit-variance column below this = constant

struct RoundArgs {
  const double* Xs;
  const float* Xf;    // FP32 copy of Xs for the Gram model (null: use Xs)
  const double* xsd;
  int n, ldx, p, p_pad, P;
  LassoParams prm;
  const double* Yn;   // n x B (ld = n) normalised expression
  const int* fold;    // n x B: 0 = not a pred cell, 1..nfolds (path variant: any non-zero)
  const int* self;    // B or null
  int mh;             // largest active set whose packed Gram block fits the shared-memory region
  int round_no;       // round index of this launch (breadcrumbs)
  LassoWs ws;
};

// ---------------------------------------------------------------------------------------------
// block reductions: every thread receives bit-identical totals (butterfly sums commute), so all
// scalar control flow below is uniform across the CTA without broadcasts.  One __syncthreads per
// call; `buf` holds 2 x 8 x 32 doubles and alternates so back-to-back calls do not race.
// ---------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void bsum(double (&v)[K], double* buf, int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* b = buf + phase * 256;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    v[k] = warp_sum(v[k]);
    if (lane == 0) b[k * 32 + warp] = v[k];
  }
  __syncthreads();
  // second stage: every group of nw lanes holds all nw partials, log2(nw) butterfly levels
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = b[k * 32 + (lane & (nw - 1))];
    for (int o = nw >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    v[k] = t;
  }
  phase ^= 1;
}

__device__ __forceinline__ double bmax(double v, double* buf, int& phase) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* b = buf + phase * 256;
  v = warp_max(v);
  if (lane == 0) b[warp] = v;
  __syncthreads();
  v = warp_max(lane < nw ? b[lane] : -INFINITY);
  phase ^= 1;
  return v;
}

__device__ __forceinline__ double2 ldg2(const double* p) {
  return __ldg(reinterpret_cast<const double2*>(p));
}

__device__ __forceinline__ void cpa16(void* smem, const void* gmem) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(a), "l"(gmem));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// ---------------------------------------------------------------------------------------------
// setup: per problem, observation weights (into the R panel, for the moment GEMMs) and the
// null-model constants of fishnet1.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lasso_setup_kernel(RoundArgs A) {
  __shared__ double red[512];
  int phase = 0;
  const int pid = blockIdx.x, tid = threadIdx.x;
  const int NF = A.prm.NF, g = pid / NF, fit = pid % NF, n = A.n;
  const double* y = A.Yn + (size_t)g * n;
  const int* fo = A.fold + (size_t)g * n;
  double cnt = 0, cnth = 0, sy = 0, lo = INFINITY, hi = -INFINITY;
  for (int i = tid; i < n; i += 256) {
    const int f = fo[i];
    const double yi = y[i];
    lo = fmin(lo, yi);
    hi = fmax(hi, yi);
    const bool pred = f != 0;
    if (pred && f != fit) { cnt += 1; sy += yi; }
    if (pred && fit > 0 && f == fit) cnth += 1;
  }
  double v3[3] = {cnt, cnth, sy};
  bsum<3>(v3, red, phase);
  lo = -bmax(-lo, red, phase);
  hi = bmax(hi, red, phase);
  const int mtrain = (int)v3[0];
  const double qv = mtrain > 0 ? 1.0 / mtrain : 0.0;
  double yb = 0, tly = 0;
  double* q = A.ws.R + (size_t)pid * A.ldx;
  for (int i = tid; i < A.ldx; i += 256) {
    double qi = 0;
    if (i < n) {
      const int f = fo[i];
      if (f != 0 && f != fit) {
        qi = qv;
        const double t = qv * y[i];
        yb += t;
        if (t > 0) tly += t * log(y[i]);
      }
    }
    q[i] = qi;
  }
  double v2[2] = {yb, tly};
  bsum<2>(v2, red, phase);
  yb = v2[0];
  tly = v2[1];
  if (tid == 0) {
    LassoProb pr;
    pr.gene = g; pr.fit = fit; pr.state = ST_INIT; pr.lidx = 0; pr.lmu = 0; pr.nin = 0; pr.nS = 0;
    pr.jerr = 0; pr.nlp = 0; pr.ngrad = 0; pr.mtrain = mtrain; pr.slot = pid;
    pr.al = 0; pr.al0 = 0; pr.qv = qv; pr.yb = yb;
    pr.az = 0; pr.dv0 = 0; pr.dvr = 0; pr.shr = 0; pr.v0 = yb; pr.sumr = 0;
    pr.cur_dev = 0; pr.cur_cv = 0; pr.ncv = v3[1];
    bool neg = lo < 0.0;
    if (mtrain < 1 || !(yb > 0.0) || neg) {
      pr.jerr = neg ? 8888 : 9999;  // fishnet: negative response / no positive response
      pr.state = ST_DONE;
    } else {
      pr.az = log(yb);
      pr.dv0 = yb * (pr.az - 1.0);
      pr.dvr = -yb + tly - pr.dv0;
      pr.shr = A.prm.thr * pr.dvr;
      if (!(pr.dvr > 0.0)) { pr.jerr = 7777; pr.state = ST_DONE; }  // constant response in this fit
    }
    A.ws.prob[pid] = pr;
    if (fit == 0) {
      A.ws.gstat[g] = (lo == hi) ? 1 : 0;  // sd(y) == 0 over ALL cells (R/expr_predict.R:37)
      A.ws.gmean[g] = mtrain > 0 ? v3[2] / mtrain : 0.0;
    }
  }
}

// XM (= Xs^T q) and XI (= (Xs^2)^T q) -> per-fit mean and 1/sd; eligibility (glmnet ju + the
// gene's own column, R/calc_estimate.R:103-110); clears the strong/active marks.
__global__ void __launch_bounds__(256) lasso_moments_kernel(RoundArgs A) {
  const int j = blockIdx.x * 256 + threadIdx.x, pid = blockIdx.y;
  if (j >= A.p_pad) return;
  const size_t o = (size_t)pid * A.p_pad + j;
  const int g = pid / A.prm.NF;
  double xi = 0.0;
  if (j < A.p) {
    const double m = A.ws.XM[o];
    const double s2 = A.ws.XI[o] - m * m;
    const int self = A.self ? A.self[g] : -1;
    if (A.xsd[j] > 0.0 && j != self && s2 > kVarTiny) xi = 1.0 / sqrt(s2);
  }
  A.ws.XI[o] = xi;
  A.ws.MM[o] = 0;
}

// null-model working weights / residual (fishnet1 prologue) -> W, WR and the GEMM panel R
__global__ void __launch_bounds__(256) lasso_resid0_kernel(RoundArgs A) {
  __shared__ double red[512];
  int phase = 0;
  const int pid = blockIdx.x, tid = threadIdx.x, n = A.n, ldx = A.ldx;
  LassoProb* PR = A.ws.prob + pid;
  const int g = PR->gene, fit = PR->fit;
  const bool dead = PR->state == ST_DONE;
  const double qv = PR->qv, yb = PR->yb, az = PR->az;
  const double* y = A.Yn + (size_t)g * n;
  const int* fo = A.fold + (size_t)g * n;
  double* w = A.ws.W + (size_t)pid * ldx;
  double* wr = A.ws.WR + (size_t)pid * ldx;
  double* r = A.ws.R + (size_t)pid * ldx;
  double sr = 0, tf = 0;
  for (int i = tid; i < ldx; i += 256) {
    double wi = 0, ri = 0;
    if (i < n && !dead) {
      const int f = fo[i];
      if (f != 0 && f != fit) {
        const double t = qv * y[i];
        wi = qv * yb;
        ri = t - wi;
        tf += t * az;
      }
    }
    w[i] = wi;
    wr[i] = ri;
    r[i] = ri;
    sr += ri;
  }
  double v2[2] = {sr, tf};
  bsum<2>(v2, red, phase);
  if (tid == 0) {
    PR->sumr = v2[0];
    if (!dead) {
      PR->cur_dev = (v2[1] - PR->v0 - PR->dv0) / PR->dvr;
      const int pos = atomicAdd(A.ws.counters, 1);
      A.ws.list0[pos] = pid;
    }
  }
}

// glmnet's automatic lambda sequence of the master fit (one CTA per gene); the folds run the same
// values as a user sequence (cv.glmnet passes lambda = master$lambda, whose first entry has been
// replaced by exp(2 log l2 - log l3), glmnet's fix.lam).
__global__ void __launch_bounds__(256) lasso_lambda_kernel(RoundArgs A) {
  __shared__ double red[512];
  int phase = 0;
  const int g = blockIdx.x, tid = threadIdx.x, pid = g * A.prm.NF;
  const LassoProb* PR = A.ws.prob + pid;
  const double sumr = PR->sumr;
  const double* Gc = A.ws.G + (size_t)PR->slot * A.p_pad;
  const double* XMc = A.ws.XM + (size_t)pid * A.p_pad;
  const double* XIc = A.ws.XI + (size_t)pid * A.p_pad;
  double m = 0.0;
  for (int j = tid; j < A.p; j += 256) {
    const double xi = XIc[j];
    if (xi != 0.0) m = fmax(m, fabs((Gc[j] - XMc[j] * sumr) * xi));
  }
  m = bmax(m, red, phase);
  if (tid == 0) {
    const int self = A.self ? A.self[g] : -1;
    const int nvars = A.p - (self >= 0 ? 1 : 0);
    const double alf = PR->mtrain < nvars ? A.prm.alf_wide : A.prm.alf_tall;
    double* lam = A.ws.lam + (size_t)g * kNLam;
    const int nlam = A.prm.nlam;
    lam[0] = kBig;
    double al = m;
    for (int l = 1; l < nlam; ++l) {
      al *= alf;
      lam[l] = al;
    }
    if (nlam > 2) lam[0] = exp(2.0 * log(lam[1]) - log(lam[2]));
    A.ws.nlam_g[g] = nlam;
  }
}

// Folds follow their gene's master fit: they stop at its length and never run more than one lambda
// ahead of its committed path.  Both decisions read this snapshot, taken between rounds, and never
// the master's live record (master and folds run in the same launch; reading the live record made
// the amount of work -- though not the result -- depend on CTA scheduling).
__global__ void lasso_snapshot_kernel(RoundArgs A) {
  const int g = blockIdx.x * 256 + threadIdx.x;
  if (g >= A.prm.B) return;
  const LassoProb& m = A.ws.prob[(size_t)g * A.prm.NF];
  A.ws.msnap[2 * g] = m.lmu;
  A.ws.msnap[2 * g + 1] = m.state == ST_DONE ? m.lmu : -1;
}

// ---------------------------------------------------------------------------------------------
// the round kernel
// ---------------------------------------------------------------------------------------------
constexpr int kMeta = 64;  // strong-set metadata staged per chunk
constexpr int kChordKeep = 4;   // IRLS steps of one launch that may reuse the launch's first Gram block
constexpr int kPart = 512;        // doubles: partial sums of the row-split streaming work items
constexpr int kHSlots = 148 * 4;  // Gram scratch slots (>= resident CTAs of the gram kernel)
constexpr int kGramSmem = kMeta * 8 + 3 * kNX * 8 + 64 + kNX * 4 - kNX * 8 + kMeta * 4 + kPart * 8;  // mg, gq, gq2, amat, gsh, ord, mamb, part; as_ lives in global memory
constexpr int kFixedSmem = 512 * 8 + 6 * kNX * 8 + kNX * 4 + 64 * 4 + kMeta * (8 + 8 + 4 + 4) + 16;

template <int T>
__device__ __forceinline__ int rebuild_strong(const int* MMc, int* SIc, int p, int* ish) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  int base = 0;
  for (int j0 = 0; j0 < p; j0 += T) {
    const int j = j0 + tid;
    const bool flag = j < p && MMc[j] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) ish[warp] = __popc(b);
    __syncthreads();
    int woff = 0, tot = 0;
    for (int w = 0; w < nw; ++w) {
      const int c = ish[w];
      if (w < warp) woff += c;
      tot += c;
    }
    if (flag) SIc[base + woff + __popc(b & ((1u << lane) - 1u))] = j;
    base += tot;
    __syncthreads();
  }
  return base;
}

// STAGED: working weights / residual live in shared memory.  XBUF: design columns are streamed
// through a two-deep cp.async ring in shared memory, so the L2 latency of column l+1 hides behind
// the reduction of column l and the axpy re-reads the column from shared memory.
template <int T, bool STAGED, bool XBUF>
__global__ void __launch_bounds__(T)
lasso_round_kernel(RoundArgs A, const int* __restrict__ list, int* __restrict__ nextlist,
                   int* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* red = reinterpret_cast<double*>(smraw);
  double* aact = red + 512;
  double* as_ = aact + kNX;
  double* va = as_ + kNX;
  double* swa = va + kNX;
  double* xma = swa + kNX;
  double* xia = xma + kNX;
  int* ia = reinterpret_cast<int*>(xia + kNX);
  int* ish = ia + kNX;
  double* mxm = reinterpret_cast<double*>(ish + 64);   // staged strong-set metadata
  double* mxi = mxm + kMeta;
  int* mk = reinterpret_cast<int*>(mxi + kMeta);
  int* mpos = mk + kMeta + 4;  // mk has one extra slot: first column of the next chunk
  double* wv = reinterpret_cast<double*>(mpos + kMeta);
  double* wrv = wv + A.ldx;
  double* xb0 = wrv + A.ldx;  // XBUF only
  double* xb1 = xb0 + A.ldx;

  const int tid = threadIdx.x;
  int phase = 0;
  unsigned ncols = 0;  // design columns this CTA streamed (roofline accounting)
  const long long t_start = clock64();
  const int pid = list[blockIdx.x];
  LassoProb* PR = A.ws.prob + pid;
  const int g = PR->gene, fit = PR->fit, slot = PR->slot;
  int state = PR->state, lidx = PR->lidx, lmu = PR->lmu, nin = PR->nin, nS = PR->nS;
  int jerr = PR->jerr, nlp = PR->nlp;
  double al = PR->al, al0 = PR->al0, az = PR->az, v0 = PR->v0, sumr = PR->sumr;
  double cur_dev = PR->cur_dev, cur_cv = PR->cur_cv;
  const double qv = PR->qv, dv0 = PR->dv0, dvr = PR->dvr, shr = PR->shr, ncv = PR->ncv;
  const int n = A.n, ldx = A.ldx, p = A.p, p_pad = A.p_pad;
  const double* __restrict__ Xs = A.Xs;
  const float* __restrict__ Xf = A.Xf;
  const double* Gc = A.ws.G + (size_t)slot * p_pad;
  const double* XMc = A.ws.XM + (size_t)pid * p_pad;
  const double* XIc = A.ws.XI + (size_t)pid * p_pad;
  int* MMc = A.ws.MM + (size_t)pid * p_pad;
  int* SIc = A.ws.SIDX + (size_t)pid * p_pad;
  int* IAg = A.ws.IA + (size_t)pid * kNX;
  double* AAg = A.ws.AACT + (size_t)pid * kNX;
  double* Wg = A.ws.W + (size_t)pid * ldx;
  double* WRg = A.ws.WR + (size_t)pid * ldx;
  const double* y = A.Yn + (size_t)g * n;
  const int* fo = A.fold + (size_t)g * n;
  const int self = A.self ? A.self[g] : -1;
  const int nvars = p - (self >= 0 ? 1 : 0);
  int nx = 2 * A.prm.dfmax + 20;
  if (nx > nvars) nx = nvars;
  if (nx > kNX) nx = kNX;
  const int ne = A.prm.dfmax;
  const int nlam = A.ws.nlam_g[g];
  const double* lamg = A.ws.lam + (size_t)g * kNLam;
  const bool master_auto = A.prm.auto_lambda && fit == 0;
  const bool is_fold = fit > 0;
  const int maxit = A.prm.maxit;
  const double fmax_ = log(DBL_MAX * 0.1);

  if (!STAGED) {
    wv = Wg;
    wrv = WRg;
  } else {
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      *reinterpret_cast<double2*>(wv + i) = *reinterpret_cast<const double2*>(Wg + i);
      *reinterpret_cast<double2*>(wrv + i) = *reinterpret_cast<const double2*>(WRg + i);
    }
  }
  for (int l = tid; l < nin; l += T) {
    const int k = IAg[l];
    ia[l] = k;
    aact[l] = AAg[l];
    xma[l] = XMc[k];
    xia[l] = XIc[k];
  }
  __syncthreads();

  // ---------------- phase A: consume the gradient of the previous round ----------------
  auto mark_pass = [&](double thresh) -> int {
    int cnt = 0;
    for (int j = tid; j < p; j += T) {
      const double xi = XIc[j];
      if (xi != 0.0 && MMc[j] == 0) {
        const double ga = fabs((Gc[j] - XMc[j] * sumr) * xi);
        if (ga > thresh) {
          MMc[j] = -1;
          ++cnt;
        }
      }
    }
    return __syncthreads_or(cnt);
  };

  bool strong_changed = false;
  if (state == ST_INIT && (nlam < 1 || A.ws.prob[(size_t)g * A.prm.NF].jerr > 0)) {
    state = ST_DONE;  // the gene's master fit failed: the whole gene falls back to the mean
  } else if (state == ST_INIT) {
    if (master_auto) {
      double m = 0.0;
      for (int j = tid; j < p; j += T) {
        const double xi = XIc[j];
        if (xi != 0.0) m = fmax(m, fabs((Gc[j] - XMc[j] * sumr) * xi));
      }
      m = bmax(m, red, phase);
      if (tid == 0) {  // the null model is solution 0 (lambda = "big")
        A.ws.a0p[(size_t)g * kNLam] = az;
        A.ws.devp[(size_t)g * kNLam] = cur_dev;
        A.ws.ninp[(size_t)g * kNLam] = 0;
      }
      lmu = 1;
      lidx = 1;
      al0 = m;
      al = lamg[1];
      if (nlam < 2) state = ST_DONE;
    } else {
      lidx = 0;
      al0 = 0.0;
      al = lamg[0];
    }
    if (state != ST_DONE) {
      strong_changed = mark_pass(2.0 * al - al0) != 0;
      state = ST_SOLVE;
    }
  } else if (state == ST_WAIT_KKT || state == ST_HOLD) {
    if (state == ST_WAIT_KKT && mark_pass(al)) {
      strong_changed = true;
      state = ST_SOLVE;  // KKT violated outside the strong set: re-solve this lambda
    } else {
      bool stop = false;
      if (state == ST_WAIT_KKT) {
      // ---- commit solution lidx ----
      if (fit == 0) {
        const size_t o = (size_t)g * kNLam + lidx;
        if (tid == 0) {
          A.ws.a0p[o] = az;
          A.ws.devp[o] = cur_dev;
          A.ws.ninp[o] = nin;
          A.ws.nlpp[o] = nlp;
        }
        double* cp = A.ws.cap + o * kNX;
        for (int l = tid; l < nin; l += T) cp[l] = aact[l];
      } else if (tid == 0) {
        A.ws.cvdev[((size_t)g * kNLam + lidx) * kMaxFolds + (fit - 1)] = cur_cv / ncv;
      }
      lmu = lidx + 1;
      if (master_auto) {
        const int mnl = kMnLam < nlam ? kMnLam : nlam;
        if (lidx + 1 >= mnl) {
          int me = 0;
          for (int l = tid; l < nin; l += T) me += aact[l] != 0.0;
          double mv[1] = {(double)me};
          bsum<1>(mv, red, phase);
          const double dprev = A.ws.devp[(size_t)g * kNLam + lidx - mnl + 1];
          if (mv[0] > ne) stop = true;
          else if ((cur_dev - dprev) / cur_dev < kSml) stop = true;
          else if (cur_dev > kDevMax) stop = true;
        }
      }
      }
      int lim = nlam;
      bool hold = false;
      if (is_fold && A.prm.auto_lambda) {
        const int mlmu = A.ws.msnap[2 * g], mdone = A.ws.msnap[2 * g + 1];
        if (mdone >= 0) { if (mdone < lim) lim = mdone; }
        else if (lidx + 1 > mlmu + 1) hold = true;  // at most one lambda ahead of the master's path
      }
      if (stop || lidx + 1 >= lim) {
        state = ST_DONE;
      } else if (hold) {
        state = ST_HOLD;
      } else {
        ++lidx;
        al0 = al;
        al = lamg[lidx];
        strong_changed = mark_pass(2.0 * al - al0) != 0;
        state = ST_SOLVE;
      }
    }
  }
  if (strong_changed) nS = rebuild_strong<T>(MMc, SIc, p, ish);

  // ---------------- phase B: IRLS + coordinate descent on the strong set ----------------
  auto intercept_update = [&](double& dlx) {
    double s[1] = {0.0};
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      const double2 r = *reinterpret_cast<const double2*>(wrv + i);
      s[0] += r.x + r.y;
    }
    bsum<1>(s, red, phase);
    const double d = s[0] / v0;
    az += d;
    dlx = fmax(dlx, v0 * d * d);
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      double2 r = *reinterpret_cast<const double2*>(wrv + i);
      const double2 w = *reinterpret_cast<const double2*>(wv + i);
      r.x -= d * w.x;
      r.y -= d * w.y;
      *reinterpret_cast<double2*>(wrv + i) = r;
    }
    sumr = s[0] - d * v0;
  };
  // column access: the thread that copies element i is the only one that reads it back, so a
  // cp.async.wait_group is all the synchronisation the ring needs
  auto col_issue = [&](int k, double* xb) {
    const double* col = Xs + (size_t)k * ldx;
    if (XBUF) {
      for (int i = 2 * tid; i < ldx; i += 2 * T) cpa16(xb + i, col + i);
      cpa_commit();
    } else if ((tid & 7) == 0) {
      for (int i = 2 * tid; i < ldx; i += 2 * T) prefetch_l1(col + i);
    }
  };
  auto ldx2 = [&](const double* base, int i) -> double2 {
    if (XBUF) return *reinterpret_cast<const double2*>(base + i);
    return ldg2(base + i);
  };
  auto axpy = [&](const double* xc, double d, double xm, double xi) {
    const double c = d * xi;
#pragma unroll 2
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      const double2 x = ldx2(xc, i);
      double2 r = *reinterpret_cast<const double2*>(wrv + i);
      const double2 w = *reinterpret_cast<const double2*>(wv + i);
      r.x -= c * w.x * (x.x - xm);
      r.y -= c * w.y * (x.y - xm);
      *reinterpret_cast<double2*>(wrv + i) = r;
    }
  };

  if (state == ST_SOLVE) {
    int irls_it = 0;
    bool big_step = false;
    while (true) {  // IRLS iterations
      const double az0 = az;
      for (int l = tid; l < nin; l += T) as_[l] = aact[l];
      bool overflow = false, maxed = false, nanfail = false;
      while (true) {  // sweeps over the strong set until one changes nothing
        ++nlp;
        double dlx = 0.0;
        ncols += nS;
        int buf = 0;
        if (nS > 0) col_issue(SIc[0], xb0);
        for (int s0 = 0; s0 < nS && !overflow && !nanfail; s0 += kMeta) {
          const int cn = nS - s0 < kMeta ? nS - s0 : kMeta;
          __syncthreads();  // previous chunk's metadata fully consumed
          if (tid < cn) {
            const int k = SIc[s0 + tid];
            mk[tid] = k;
            mxm[tid] = XMc[k];
            mxi[tid] = XIc[k];
            mpos[tid] = MMc[k];
          } else if (tid == cn && s0 + cn < nS) {
            mk[cn] = SIc[s0 + cn];  // first column of the next chunk, for the prefetch
          }
          __syncthreads();
          for (int j = 0; j < cn; ++j) {
            const int k = mk[j];
            double* xcur = buf ? xb1 : xb0;
            if (s0 + j + 1 < nS) {
              col_issue(mk[j + 1], buf ? xb0 : xb1);
              if (XBUF) cpa_wait<1>();
            } else if (XBUF) {
              cpa_wait<0>();
            }
            const double* xc = XBUF ? xcur : Xs + (size_t)k * ldx;
            buf ^= 1;
            int pos = mpos[j];
            const double xm = mxm[j], xi = mxi[j];
            const double ak = pos > 0 ? aact[pos - 1] : 0.0;  // read before the reduction's barrier
            double d3[3] = {0.0, 0.0, 0.0}, e3[3] = {0.0, 0.0, 0.0};
#pragma unroll 2
            for (int i = 2 * tid; i < ldx; i += 2 * T) {
              const double2 x = ldx2(xc, i);
              const double2 r = *reinterpret_cast<const double2*>(wrv + i);
              const double2 w = *reinterpret_cast<const double2*>(wv + i);
              d3[0] += r.x * x.x;
              e3[0] += r.y * x.y;
              const double wx0 = w.x * x.x, wx1 = w.y * x.y;
              d3[1] += wx0;
              e3[1] += wx1;
              d3[2] += wx0 * x.x;
              e3[2] += wx1 * x.y;
            }
            d3[0] += e3[0]; d3[1] += e3[1]; d3[2] += e3[2];
            bsum<3>(d3, red, phase);
            const double sw = xi * (d3[1] - xm * v0);
            const double v = xi * xi * (d3[2] - 2.0 * xm * d3[1] + xm * xm * v0);
            const double gk = xi * (d3[0] - xm * sumr);
            const double u = gk + v * ak;
            const double au = fabs(u) - al;
            const double anew = au > 0.0 ? copysign(au, u) / v : 0.0;
            if (pos > 0 && tid == 0) {
              va[pos - 1] = v;
              swa[pos - 1] = sw;
            }
            if (anew == ak) continue;
            if (anew != anew) { nanfail = true; break; }
            const double d = anew - ak;
            dlx = fmax(dlx, v * d * d);
            axpy(xc, d, xm, xi);
            sumr -= d * sw;
            if (pos < 0) {
              ++nin;
              if (nin > nx) { overflow = true; break; }
              pos = nin;
              if (tid == 0) {
                MMc[k] = nin;
                ia[nin - 1] = k;
                xma[nin - 1] = xm;
                xia[nin - 1] = xi;
                va[nin - 1] = v;
                swa[nin - 1] = sw;
                as_[nin - 1] = 0.0;
              }
            }
            if (tid == 0) aact[pos - 1] = anew;
            // aact / va are next read after at least one __syncthreads (the next bsum)
          }
        }
        if (XBUF) cpa_wait<0>();
        if (overflow || nanfail) break;
        intercept_update(dlx);
        if (dlx < shr) break;
        if (!(dlx == dlx)) { nanfail = true; break; }
        if (nlp > maxit) { maxed = true; break; }
        while (true) {  // sweeps over the ever-active list
          ++nlp;
          dlx = 0.0;
          ncols += nin;
          buf = 0;
          if (nin > 0) col_issue(ia[0], xb0);
          for (int l = 0; l < nin; ++l) {
            const int k = ia[l];
            double* xcur = buf ? xb1 : xb0;
            if (l + 1 < nin) {
              col_issue(ia[l + 1], buf ? xb0 : xb1);
              if (XBUF) cpa_wait<1>();
            } else if (XBUF) {
              cpa_wait<0>();
            }
            const double* xc = XBUF ? xcur : Xs + (size_t)k * ldx;
            buf ^= 1;
            const double xm = xma[l], xi = xia[l], v = va[l], ak = aact[l], sw = swa[l];
            double d1[1] = {0.0}, e1 = 0.0;
#pragma unroll 2
            for (int i = 2 * tid; i < ldx; i += 2 * T) {
              const double2 x = ldx2(xc, i);
              const double2 r = *reinterpret_cast<const double2*>(wrv + i);
              d1[0] += r.x * x.x;
              e1 += r.y * x.y;
            }
            d1[0] += e1;
            bsum<1>(d1, red, phase);
            const double gk = xi * (d1[0] - xm * sumr);
            const double u = gk + v * ak;
            const double au = fabs(u) - al;
            const double anew = au > 0.0 ? copysign(au, u) / v : 0.0;
            if (anew == ak) continue;
            if (anew != anew) { nanfail = true; break; }
            const double d = anew - ak;
            dlx = fmax(dlx, v * d * d);
            axpy(xc, d, xm, xi);
            sumr -= d * sw;
            if (tid == 0) aact[l] = anew;  // all reads of aact[l] precede the reduction's barrier
          }
          if (XBUF) cpa_wait<0>();
          if (nanfail) break;
          intercept_update(dlx);
          if (dlx < shr) break;
          if (!(dlx == dlx)) { nanfail = true; break; }
          if (nlp > maxit) { maxed = true; break; }
        }
        if (maxed || nanfail) break;
      }
      if (overflow) { jerr = -10000 - (lidx + 1); state = ST_DONE; break; }
      if (maxed) { jerr = -(lidx + 1); state = ST_DONE; break; }
      if (nanfail) { jerr = 9998; state = ST_DONE; break; }

      // ---- re-linearise: f = az + sum_l a_l x~_l, w = q exp(f), wr = t - w ----
      __syncthreads();
      double c0p[1] = {0.0};
      for (int l = tid; l < nin; l += T) {
        const double b = aact[l] * xia[l];
        swa[l] = b;  // swa is dead until the next strong sweep recomputes it
        c0p[0] += b * xma[l];
      }
      bsum<1>(c0p, red, phase);
      const double c0 = az - c0p[0];
      ncols += nin;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};  // sum w, sum t f, sum wr, held-out deviance
      for (int i = 2 * tid; i < ldx; i += 2 * T) {
        double f0 = c0, f1 = c0;
#pragma unroll 4
        for (int l = 0; l < nin; ++l) {
          const double b = swa[l];
          if (b != 0.0) {
            const double2 x = ldg2(Xs + (size_t)ia[l] * ldx + i);
            f0 += b * x.x;
            f1 += b * x.y;
          }
        }
        double wn[2] = {0.0, 0.0}, rn[2] = {0.0, 0.0};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ii = i + e;
          if (ii < n) {
            const int f = fo[ii];
            const double fe = e ? f1 : f0;
            if (f != 0 && f != fit) {
              const double yi = y[ii];
              const double t = qv * yi;
              const double fc = fabs(fe) < fmax_ ? fe : copysign(fmax_, fe);
              wn[e] = qv * exp(fc);
              rn[e] = t - wn[e];
              acc[0] += wn[e];
              acc[1] += t * fe;
              acc[2] += rn[e];
            } else if (f != 0 && is_fold) {  // held-out row of this fold
              const double yi = y[ii];
              const double devy = yi == 0.0 ? 0.0 : yi * log(yi) - yi;
              acc[3] += 2.0 * (devy - (yi * fe - exp(fe)));
            }
          }
        }
        *reinterpret_cast<double2*>(wv + i) = make_double2(wn[0], wn[1]);
        *reinterpret_cast<double2*>(wrv + i) = make_double2(rn[0], rn[1]);
      }
      bsum<4>(acc, red, phase);
      v0 = acc[0];
      sumr = acc[2];
      cur_dev = (acc[1] - v0 - dv0) / dvr;
      cur_cv = acc[3];
      // ---- outer convergence (fishnet1: intercept and every active coefficient) ----
      int bad = v0 * (az - az0) * (az - az0) >= shr;
      for (int l = tid; l < nin; l += T) {
        const double d = aact[l] - as_[l];
        bad |= va[l] * d * d >= shr;
      }
      if (!__syncthreads_or(bad)) {
        state = ST_WAIT_KKT;
        break;
      }
    }
  }

  // ---------------- write back ----------------
  __syncthreads();
  if (STAGED) {
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      *reinterpret_cast<double2*>(Wg + i) = *reinterpret_cast<const double2*>(wv + i);
      *reinterpret_cast<double2*>(WRg + i) = *reinterpret_cast<const double2*>(wrv + i);
    }
  }
  for (int l = tid; l < nin && l < kNX; l += T) {
    IAg[l] = ia[l];
    AAg[l] = aact[l];
  }
  if (state != ST_DONE) {
    if (tid == 0) {
      const int pos = atomicAdd(counter, 1);
      nextlist[pos] = pid;
      ish[40] = pos;
    }
    __syncthreads();
    const int pos = ish[40];
    double* r = A.ws.R + (size_t)pos * ldx;
    for (int i = 2 * tid; i < ldx; i += 2 * T)
      *reinterpret_cast<double2*>(r + i) = *reinterpret_cast<const double2*>(wrv + i);
    if (tid == 0) PR->slot = pos;
  }
  if (tid == 0) {
    atomicAdd(A.ws.colcount, (unsigned long long)ncols);
    const unsigned long long cyc = (unsigned long long)(clock64() - t_start);
    atomicAdd(A.ws.colcount + 1, cyc);   // sum of CTA lifetimes
    atomicMax(A.ws.colcount + 2, cyc);   // longest CTA of the block of rounds (straggler)
    atomicAdd(A.ws.colcount + 3, (unsigned long long)(state == ST_WAIT_KKT || state == ST_DONE));
    PR->state = state; PR->lidx = lidx; PR->lmu = lmu; PR->nin = nin; PR->nS = nS;
    PR->jerr = jerr; PR->nlp = nlp; PR->ngrad += 1;
    PR->al = al; PR->al0 = al0; PR->az = az; PR->v0 = v0; PR->sumr = sumr;
    PR->cur_dev = cur_dev; PR->cur_cv = cur_cv;
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming helpers of the Gram kernel.  Design columns come out of L2 (~250 cycles unloaded, far
// more under load); inside the big round kernel ptxas has no registers left to keep more than one
// load per lane in flight, which leaves every streaming loop latency bound.  These are separate
// functions (own register allocation) that issue a batch of kMLP 16-byte loads per lane before the
// first use.  The per-lane accumulation order is exactly that of the plain loops they replace.
// ---------------------------------------------------------------------------------------------
constexpr int kMLP = 8;

__device__ __forceinline__ double2 ldg2_if(const double* p, bool ok) {
  return ok ? ldg2(p) : make_double2(0.0, 0.0);
}
__device__ __forceinline__ double2 lds2_if(const double* p, bool ok) {
  return ok ? *reinterpret_cast<const double2*>(p) : make_double2(0.0, 0.0);
}

// sum_i a_i z_i : z a design column (global), a an n-vector (shared or global).  Every lane runs
// the same number of batches (out-of-range slots load nothing and add an exact zero), so the whole
// column costs ceil(ldx / (64 kMLP)) exposed L2 round trips.
__device__ __noinline__ double wdot_a(const double* __restrict__ z, const double* a, int r0, int ldx) {
  const int lane = threadIdx.x & 31;
  double s0 = 0.0, s1 = 0.0;
  for (int i = r0 + 2 * lane; i < ldx; i += 64 * kMLP) {
    double2 zz[kMLP];
#pragma unroll
    for (int u = 0; u < kMLP; ++u) zz[u] = ldg2_if(z + i + 64 * u, i + 64 * u < ldx);
#pragma unroll
    for (int u = 0; u < kMLP; ++u) {
      const double2 r = lds2_if(a + i + 64 * u, i + 64 * u < ldx);
      s0 += r.x * zz[u].x;
      s1 += r.y * zz[u].y;
    }
  }
  return warp_sum(s0 + s1);
}

// (sum_i w_i z_i, sum_i r_i z_i)
__device__ __noinline__ double2 wdot_wr(const double* __restrict__ z, const double* w,
                                        const double* r, int ldx) {
  const int lane = threadIdx.x & 31;
  double pa = 0.0, pb = 0.0, ra = 0.0, rb = 0.0;
  for (int i = 2 * lane; i < ldx; i += 64 * kMLP) {
    double2 zz[kMLP];
#pragma unroll
    for (int u = 0; u < kMLP; ++u) zz[u] = ldg2_if(z + i + 64 * u, i + 64 * u < ldx);
#pragma unroll
    for (int u = 0; u < kMLP; ++u) {
      const bool ok = i + 64 * u < ldx;
      const double2 ww = lds2_if(w + i + 64 * u, ok);
      const double2 rr = lds2_if(r + i + 64 * u, ok);
      pa += ww.x * zz[u].x;
      pb += ww.y * zz[u].y;
      ra += rr.x * zz[u].x;
      rb += rr.y * zz[u].y;
    }
  }
  return make_double2(warp_sum(pa + pb), warp_sum(ra + rb));
}

// sum_i w_i x_i z_i : x the entering column (shared by all warps, L1), z a member column
__device__ __noinline__ double wdot_wxz(const double* __restrict__ z, const double* __restrict__ x,
                                        const double* w, int ldx) {
  const int lane = threadIdx.x & 31;
  double s0 = 0.0, s1 = 0.0;
  constexpr int U = kMLP / 2;
  for (int i = 2 * lane; i < ldx; i += 64 * U) {
    double2 zz[U], xx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      zz[u] = ldg2_if(z + i + 64 * u, i + 64 * u < ldx);
      xx[u] = ldg2_if(x + i + 64 * u, i + 64 * u < ldx);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double2 ww = lds2_if(w + i + 64 * u, i + 64 * u < ldx);
      s0 += ww.x * xx[u].x * zz[u].x;
      s1 += ww.y * xx[u].y * zz[u].y;
    }
  }
  return warp_sum(s0 + s1);
}

// Four columns per call: the loads of all four are in flight together (4 x kMLP4 16-byte loads per
// lane) and the shared-memory operand is read once.  Per-column accumulation order is unchanged.
constexpr int kMLP4 = 4;

// out[c] = sum_i a_i z_c,i
__device__ __noinline__ void wdot_a4(const double* __restrict__ z0, const double* __restrict__ z1,
                                     const double* __restrict__ z2, const double* __restrict__ z3,
                                     const double* a, int ldx, double* out) {
  const int lane = threadIdx.x & 31;
  const double* zp[4] = {z0, z1, z2, z3};
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  for (int i = 2 * lane; i < ldx; i += 64 * kMLP4) {
    double2 zz[4][kMLP4];
#pragma unroll
    for (int u = 0; u < kMLP4; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) zz[c][u] = ldg2_if(zp[c] + i + 64 * u, i + 64 * u < ldx);
#pragma unroll
    for (int u = 0; u < kMLP4; ++u) {
      const double2 r = lds2_if(a + i + 64 * u, i + 64 * u < ldx);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s0[c] += r.x * zz[c][u].x;
        s1[c] += r.y * zz[c][u].y;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = warp_sum(s0[c] + s1[c]);
}

// out[c] = sum_i w_i x_i z_c,i
__device__ __noinline__ void wdot_wxz4(const double* __restrict__ z0, const double* __restrict__ z1,
                                       const double* __restrict__ z2, const double* __restrict__ z3,
                                       const double* __restrict__ x, const double* w, int ldx,
                                       double* out) {
  const int lane = threadIdx.x & 31;
  const double* zp[4] = {z0, z1, z2, z3};
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  constexpr int U = 3;
  for (int i = 2 * lane; i < ldx; i += 64 * U) {
    double2 zz[4][U], xx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = i + 64 * u < ldx;
#pragma unroll
      for (int c = 0; c < 4; ++c) zz[c][u] = ldg2_if(zp[c] + i + 64 * u, ok);
      xx[u] = ldg2_if(x + i + 64 * u, ok);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double2 ww = lds2_if(w + i + 64 * u, i + 64 * u < ldx);
      const double wx0 = ww.x * xx[u].x, wx1 = ww.y * xx[u].y;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s0[c] += wx0 * zz[c][u].x;
        s1[c] += wx1 * zz[c][u].y;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = warp_sum(s0[c] + s1[c]);
}

__device__ __forceinline__ float4 ldg4f_if(const float* p, bool ok) {
  return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// wdot_wxz4 on the FP32 copy of the design (curvature model only): four rows per 16-byte load, the
// products and sums in FP64
__device__ __noinline__ void wdot_wxz4f(const float* __restrict__ z0, const float* __restrict__ z1,
                                        const float* __restrict__ z2, const float* __restrict__ z3,
                                        const float* __restrict__ x, const double* w, int r0,
                                        int ldx, double* out) {
  const int lane = threadIdx.x & 31;
  const float* zp[4] = {z0, z1, z2, z3};
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  constexpr int U = 3;
  for (int i = r0 + 4 * lane; i < ldx; i += 128 * U) {
    float4 zz[4][U], xx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = i + 128 * u < ldx;
#pragma unroll
      for (int c = 0; c < 4; ++c) zz[c][u] = ldg4f_if(zp[c] + i + 128 * u, ok);
      xx[u] = ldg4f_if(x + i + 128 * u, ok);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = i + 128 * u < ldx;
      const double2 wa = lds2_if(w + i + 128 * u, ok), wb = lds2_if(w + i + 128 * u + 2, ok);
      const double wx0 = wa.x * (double)xx[u].x, wx1 = wa.y * (double)xx[u].y;
      const double wx2 = wb.x * (double)xx[u].z, wx3 = wb.y * (double)xx[u].w;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s0[c] += wx0 * (double)zz[c][u].x;
        s1[c] += wx1 * (double)zz[c][u].y;
        s0[c] += wx2 * (double)zz[c][u].z;
        s1[c] += wx3 * (double)zz[c][u].w;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = warp_sum(s0[c] + s1[c]);
}

// Screening dots on the FP32 copy: out[c] = sum_i a_i zf_c,i and outabs[c] = sum_i |a_i zf_c,i|.
// Every FP32 entry is its FP64 original times (1 + e), |e| <= 2^-24, so the exact dot lies within
// 2^-24 * outabs[c] of out[c]: candidates that this interval separates from the threshold need no
// FP64 pass (the strong-set check), the others are recomputed exactly.
__device__ __noinline__ void wdot_a4f(const float* __restrict__ z0, const float* __restrict__ z1,
                                      const float* __restrict__ z2, const float* __restrict__ z3,
                                      const double* a, int r0, int ldx, double* out,
                                      double* outabs) {
  const int lane = threadIdx.x & 31;
  const float* zp[4] = {z0, z1, z2, z3};
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0}, t0[4] = {0, 0, 0, 0}, t1[4] = {0, 0, 0, 0};
  constexpr int U = 3;
  for (int i = r0 + 4 * lane; i < ldx; i += 128 * U) {
    float4 zz[4][U];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) zz[c][u] = ldg4f_if(zp[c] + i + 128 * u, i + 128 * u < ldx);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = i + 128 * u < ldx;
      const double2 ra = lds2_if(a + i + 128 * u, ok), rb = lds2_if(a + i + 128 * u + 2, ok);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double p0 = ra.x * (double)zz[c][u].x, p1 = ra.y * (double)zz[c][u].y;
        const double p2 = rb.x * (double)zz[c][u].z, p3 = rb.y * (double)zz[c][u].w;
        s0[c] += p0;
        s1[c] += p1;
        s0[c] += p2;
        s1[c] += p3;
        t0[c] += fabs(p0);
        t1[c] += fabs(p1);
        t0[c] += fabs(p2);
        t1[c] += fabs(p3);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    out[c] = warp_sum(s0[c] + s1[c]);
    outabs[c] = warp_sum(t0[c] + t1[c]);
  }
}

// (pout[c], rout[c]) = (sum_i w_i z_c,i, sum_i r_i z_c,i)
__device__ __noinline__ void wdot_wr4(const double* __restrict__ z0, const double* __restrict__ z1,
                                      const double* __restrict__ z2, const double* __restrict__ z3,
                                      const double* w, const double* r, int r0, int ldx,
                                      double* pout, double* rout) {
  const int lane = threadIdx.x & 31;
  const double* zp[4] = {z0, z1, z2, z3};
  double pa[4] = {0, 0, 0, 0}, pb[4] = {0, 0, 0, 0}, ra[4] = {0, 0, 0, 0}, rb[4] = {0, 0, 0, 0};
  constexpr int U = 3;
  for (int i = r0 + 2 * lane; i < ldx; i += 64 * U) {
    double2 zz[4][U];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) zz[c][u] = ldg2_if(zp[c] + i + 64 * u, i + 64 * u < ldx);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = i + 64 * u < ldx;
      const double2 ww = lds2_if(w + i + 64 * u, ok);
      const double2 rr = lds2_if(r + i + 64 * u, ok);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        pa[c] += ww.x * zz[c][u].x;
        pb[c] += ww.y * zz[c][u].y;
        ra[c] += rr.x * zz[c][u].x;
        rb[c] += rr.y * zz[c][u].y;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    pout[c] = warp_sum(pa[c] + pb[c]);
    rout[c] = warp_sum(ra[c] + rb[c]);
  }
}

// (s0, s1) += sum_l c_l * Xs[i .. i+1, k_l] for one row pair; zero coefficients add an exact zero
__device__ __noinline__ double2 rows_comb(const double* __restrict__ Xi, size_t ldx, const int* ks,
                                          const double* cs, int m, double s0, double s1) {
  int l = 0;
  for (; l + kMLP <= m; l += kMLP) {
    double2 xx[kMLP];
#pragma unroll
    for (int u = 0; u < kMLP; ++u) xx[u] = ldg2(Xi + (size_t)ks[l + u] * ldx);
#pragma unroll
    for (int u = 0; u < kMLP; ++u) {
      const double c = cs[l + u];
      s0 += c * xx[u].x;
      s1 += c * xx[u].y;
    }
  }
  for (; l < m; ++l) {
    const double c = cs[l];
    const double2 x = ldg2(Xi + (size_t)ks[l] * ldx);
    s0 += c * x.x;
    s1 += c * x.y;
  }
  return make_double2(s0, s1);
}

// One Gram-space sweep executed by a single warp: q lives in registers (Q entries per lane), the
// Gram block in shared memory, packed lower-triangular by rows
// (element (r, c), r >= c, at r(r+1)/2 + c).  Column l seen by lane j is the contiguous run
// l(l+1)/2 + j for j <= l and the triangular-stride run j(j+1)/2 + l below it; triangular numbers
// visit every residue mod 16 exactly twice per 32 lanes, so both halves are bank-conflict free.
//
// A sweep is one long dependent chain (step t needs q after step t-1), and the FP64 pipe it runs
// on is shared with the DMMA / DFMA streams of the neighbouring warps, so the step is written to
// keep that chain minimal: shuffle -> DFMA (u) -> DADD (|u| - lambda) -> soft threshold in integer
// ALU ops on the sign word -> DMUL (new coefficient) -> select -> DADD (d) -> DFMA (q update).
// Nothing branches on the chain (d = 0 makes every update an exact no-op), everything else of
// step t+1 -- scalars, the Gram column, the index -- is loaded while step t's chain drains, and in
// natural order the shuffle source register is static.
struct SweepAcc {
  double dlx, q0;
  unsigned dty, bad;
};

__device__ __forceinline__ double sweep_step(double ql, double v, double ak, double rv, double al,
                                             double sw, SweepAcc& A, double& anew_out) {
  const double u = fma(v, ak, ql);
  const double r = fabs(u) - al;
  // anew = r > 0 ? sign(r, u) * rv : 0 (glmnet), the sign transfer in integer ALU ops and the
  // compare running beside the multiply instead of in front of it
  const double cs = __hiloint2double((__double2hiint(r) & 0x7fffffff) | (__double2hiint(u) & 0x80000000),
                                     __double2loint(r));
  const double t = cs * rv;
  const double anew = r > 0.0 ? t : 0.0;
  const double d = anew - ak;
  // off the chain
  A.dlx = fmax(A.dlx, v * d * d);
  A.q0 = fma(-d, sw, A.q0);
  A.bad |= anew != anew;
  A.dty |= (unsigned)(__double2hiint(d) << 1) | (unsigned)__double2loint(d);
  anew_out = anew;
  return d;
}

template <int Q, bool BYIDX>
__device__ __noinline__ void gram_sweep_warp_sm(const double* Hp, int m, const int* ord, double* gq,
                                                const double* va, const double* vinv, double* aact,
                                                const double* swa, double al, double* gsh,
                                                double dlx, double q0) {
  const int lane = threadIdx.x & 31;
  double qr[Q], hn[Q];
  int tj[Q];
#pragma unroll
  for (int e = 0; e < Q; ++e) {
    const int j = lane + 32 * e;
    qr[e] = j < m ? gq[j] : 0.0;
    tj[e] = j * (j + 1) / 2;
  }
  auto ldcol = [&](int l) {
    const int tl = l * (l + 1) / 2;
#pragma unroll
    for (int e = 0; e < Q; ++e) {
      const int j = lane + 32 * e;
      hn[e] = j < m ? Hp[j <= l ? tl + j : tj[e] + l] : 0.0;
    }
  };
  SweepAcc A = {dlx, q0, 0u, 0u};
  int ln = BYIDX ? ord[0] : 0;
  double vn = va[ln], an = aact[ln], sn = swa[ln], rn = vinv[ln];
  ldcol(ln);
  if (BYIDX) {
    for (int t = 0; t < m; ++t) {
      const int l = ln;
      const double v = vn, ak = an, sw = sn, rv = rn;
      double hc[Q];
#pragma unroll
      for (int e = 0; e < Q; ++e) hc[e] = hn[e];
      if (t + 1 < m) {
        ln = ord[t + 1];
        vn = va[ln]; an = aact[ln]; sn = swa[ln]; rn = vinv[ln];
        ldcol(ln);
      }
      double sel = qr[0];
      const int qi = l >> 5;
#pragma unroll
      for (int e = 1; e < Q; ++e) sel = qi == e ? qr[e] : sel;
      const double ql = __shfl_sync(0xffffffffu, sel, l & 31);
      double anew;
      const double d = sweep_step(ql, v, ak, rv, al, sw, A, anew);
#pragma unroll
      for (int e = 0; e < Q; ++e) qr[e] = fma(-d, hc[e], qr[e]);
      if (lane == 0) aact[l] = anew;
    }
  } else {
#pragma unroll
    for (int e0 = 0; e0 < Q; ++e0) {
      const int tend = m - 32 * e0 < 32 ? m - 32 * e0 : 32;
      for (int tt = 0; tt < tend; ++tt) {
        const int l = 32 * e0 + tt;
        const double v = vn, ak = an, sw = sn, rv = rn;
        double hc[Q];
#pragma unroll
        for (int e = 0; e < Q; ++e) hc[e] = hn[e];
        if (l + 1 < m) {
          ln = l + 1;
          vn = va[ln]; an = aact[ln]; sn = swa[ln]; rn = vinv[ln];
          ldcol(ln);
        }
        const double ql = __shfl_sync(0xffffffffu, qr[e0], tt);
        double anew;
        const double d = sweep_step(ql, v, ak, rv, al, sw, A, anew);
#pragma unroll
        for (int e = 0; e < Q; ++e) qr[e] = fma(-d, hc[e], qr[e]);
        if (lane == 0) aact[l] = anew;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < Q; ++e) {
    const int j = lane + 32 * e;
    if (j < m) gq[j] = qr[e];
  }
  if (lane == 0) {
    gsh[0] = A.bad ? NAN : A.dlx;
    gsh[1] = A.q0;
    gsh[2] = A.dty ? 1.0 : 0.0;
  }
}

template <int Q>
__device__ __forceinline__ void gram_sweep_sm(const double* Hp, int m, bool by_index, const int* ord,
                                              double* gq, const double* va, const double* vinv,
                                              double* aact, const double* swa, double al,
                                              double* gsh, double dlx, double q0) {
  if (by_index) gram_sweep_warp_sm<Q, true>(Hp, m, ord, gq, va, vinv, aact, swa, al, gsh, dlx, q0);
  else gram_sweep_warp_sm<Q, false>(Hp, m, ord, gq, va, vinv, aact, swa, al, gsh, dlx, q0);
}

// The same step sequence with the Gram block in global memory (active sets too large for the
// packed shared-memory copy): the column of step t is loaded at the top of the step and consumed
// after the scalar chain, and pulled into L1 NB steps ahead with prefetch.global.L1 (one sector
// per four lanes), so the load is an L1 hit hidden behind the chain.  Q up to 12 keeps active sets
// of up to 384 variables in one warp's registers.
template <int Q, bool BYIDX>
__device__ __noinline__ void gram_sweep_warp_g(const double* __restrict__ H, int m, const int* ord,
                                               double* gq, const double* va, const double* vinv,
                                               double* aact, const double* swa, double al,
                                               double* gsh, double dlx, double q0) {
  constexpr int NB = 4, ldh = kNX;
  const int lane = threadIdx.x & 31;
  double qr[Q];
#pragma unroll
  for (int e = 0; e < Q; ++e) {
    const int j = lane + 32 * e;
    qr[e] = j < m ? gq[j] : 0.0;
  }
  auto pf = [&](int t) {
    if (t < m && (lane & 3) == 0) {
      const double* c = H + (size_t)(BYIDX ? ord[t] : t) * ldh + lane;
#pragma unroll
      for (int e = 0; e < Q; ++e)
        if (lane + 32 * e < m) prefetch_l1(c + 32 * e);
    }
  };
#pragma unroll
  for (int b = 0; b < NB; ++b) pf(b);
  SweepAcc A = {dlx, q0, 0u, 0u};
  int ln = BYIDX ? ord[0] : 0;
  double vn = va[ln], an = aact[ln], sn = swa[ln], rn = vinv[ln];
  auto step = [&](int t, int l, double ql_src, int src_lane) {
    const double v = vn, ak = an, sw = sn, rv = rn;
    double hc[Q];
    const double* c = H + (size_t)l * ldh + lane;
#pragma unroll
    for (int e = 0; e < Q; ++e) hc[e] = lane + 32 * e < m ? c[32 * e] : 0.0;
    pf(t + NB);
    if (t + 1 < m) {
      ln = BYIDX ? ord[t + 1] : t + 1;
      vn = va[ln]; an = aact[ln]; sn = swa[ln]; rn = vinv[ln];
    }
    const double ql = __shfl_sync(0xffffffffu, ql_src, src_lane);
    double anew;
    const double d = sweep_step(ql, v, ak, rv, al, sw, A, anew);
#pragma unroll
    for (int e = 0; e < Q; ++e) qr[e] = fma(-d, hc[e], qr[e]);
    if (lane == 0) aact[l] = anew;
  };
  if (BYIDX) {
    for (int t = 0; t < m; ++t) {
      const int l = ln;
      double sel = qr[0];
      const int qi = l >> 5;
#pragma unroll
      for (int e = 1; e < Q; ++e) sel = qi == e ? qr[e] : sel;
      step(t, l, sel, l & 31);
    }
  } else {
#pragma unroll
    for (int e0 = 0; e0 < Q; ++e0) {
      const int tend = m - 32 * e0 < 32 ? m - 32 * e0 : 32;
      for (int tt = 0; tt < tend; ++tt) step(32 * e0 + tt, 32 * e0 + tt, qr[e0], tt);
    }
  }
#pragma unroll
  for (int e = 0; e < Q; ++e) {
    const int j = lane + 32 * e;
    if (j < m) gq[j] = qr[e];
  }
  if (lane == 0) {
    gsh[0] = A.bad ? NAN : A.dlx;
    gsh[1] = A.q0;
    gsh[2] = A.dty ? 1.0 : 0.0;
  }
}

template <int Q>
__device__ __forceinline__ void gram_sweep_g(const double* H, int m, bool by_index, const int* ord,
                                             double* gq, const double* va, const double* vinv,
                                             double* aact, const double* swa, double al, double* gsh,
                                             double dlx, double q0) {
  if (by_index) gram_sweep_warp_g<Q, true>(H, m, ord, gq, va, vinv, aact, swa, al, gsh, dlx, q0);
  else gram_sweep_warp_g<Q, false>(H, m, ord, gq, va, vinv, aact, swa, al, gsh, dlx, q0);
}

// The Gram-space round kernel (default): same state machine and phase A as lasso_round_kernel,
// phase B reformulated so that no coordinate update touches an n-long vector.
// MINB = resident CTAs per SM the register allocation is planned for.  With two CTAs per SM a thread
// has 128 registers and ptxas does NOT keep the 12-16 loads of a streaming helper in flight (it
// interleaves them with their uses: SASS of wdot_a4, profiles/README.md); with one CTA per SM
// (large n, where W / WR fill shared memory anyway) 255 registers let every batch issue as written.
template <int T, bool STAGED, int MINB>
__global__ void __launch_bounds__(T, MINB)
lasso_round_gram_kernel(RoundArgs A, const int* __restrict__ list, int* __restrict__ nextlist,
                   int* __restrict__ counter) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* red = reinterpret_cast<double*>(smraw);
  double* aact = red + 512;
  double* va = aact + kNX;
  double* swa = va + kNX;
  double* xma = swa + kNX;
  double* xia = xma + kNX;
  int* ia = reinterpret_cast<int*>(xia + kNX);
  int* ish = ia + kNX;
  double* mxm = reinterpret_cast<double*>(ish + 64);   // staged strong-set metadata
  double* mxi = mxm + kMeta;
  int* mk = reinterpret_cast<int*>(mxi + kMeta);
  int* mpos = mk + kMeta + 4;  // mk has one extra slot: first column of the next chunk
  double* mg = reinterpret_cast<double*>(mpos + kMeta);  // staged strong-set gradients
  double* gq = mg + kMeta;                                // gradient on the active set
  double* gq2 = gq + kNX;                                  // second buffer of the gradient
  double* amat = gq2 + kNX;                                // coefficients at the last materialisation
  double* gsh = amat + kNX;                                // 8 doubles: sweep results broadcast from warp 0
  int* ord = reinterpret_cast<int*>(gsh + 8);              // active positions by variable index
  int* mamb = ord + kNX;                                   // FP32 screening: candidates that need the FP64 dot
  double* part = reinterpret_cast<double*>(mamb + kMeta);  // partial dots of (column group, row segment) items
  double* wv = part + kPart;
  double* wrv = wv + A.ldx;
  double* Hsm = STAGED ? wrv + A.ldx : wv;                 // packed Gram block (A.mh rows at most)
  const int mh = A.mh;
  int hs_m = 0;  // rows of the shared-memory block that mirror H (usable by the sweep iff == m)

  const int tid = threadIdx.x;
  int phase = 0;
  unsigned ncols = 0;  // design columns this CTA streamed (roofline accounting)
  unsigned long long gram_fma = 0;  // FP64 FMAs issued on DMMA for Gram blocks
  long long tph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // cycles: phaseA, vectors, gramP, sweeps, materialise, inactive, relin, irls count; [8] sweep steps [9] sweeps [10] inactive columns [11] checks [12] sum m at Gram build
  long long tmark = clock64();
  const bool timing = (A.prm.dbg & 128) != 0;
  const bool mlp = (A.prm.dbg & 256) == 0;  // batched-load streaming helpers (bit 256: plain loops)
  auto lap = [&](int k) {
    if (timing) { const long long t = clock64(); tph[k] += t - tmark; tmark = t; }
  };
  int ninact = 0;
  int* crumb_p = (A.prm.dbg & 4096) && A.ws.d_crumb ? A.ws.d_crumb + (size_t)(blockIdx.x % kCrumbSlots) * 8 : nullptr;
  // The list is in completion order of the previous round (quick problems first); taking it from
  // the back starts the long-running problems first, so the round does not end on stragglers.
  const int lpos = (A.prm.dbg & 2048) ? (int)blockIdx.x : (int)(gridDim.x - 1 - blockIdx.x);
  if (list[lpos] < 0) return;  // hole of a gene-aligned list (a fit of the gene that is already done)
  int* INc = A.ws.INACT + (size_t)list[lpos] * A.p_pad;
  const long long t_start = clock64();
  const int pid = list[lpos];
  LassoProb* PR = A.ws.prob + pid;
  const int g = PR->gene, fit = PR->fit, slot = PR->slot;
  int state = PR->state, lidx = PR->lidx, lmu = PR->lmu, nin = PR->nin, nS = PR->nS;
  int jerr = PR->jerr, nlp = PR->nlp;
  double al = PR->al, al0 = PR->al0, az = PR->az, v0 = PR->v0, sumr = PR->sumr;
  double cur_dev = PR->cur_dev, cur_cv = PR->cur_cv;
  const double qv = PR->qv, dv0 = PR->dv0, dvr = PR->dvr, shr = PR->shr, ncv = PR->ncv;
  const double shr_in = shr * A.prm.inner_scale;  // exit test of the Gram-space sweeps
  const int n = A.n, ldx = A.ldx, p = A.p, p_pad = A.p_pad;
  const double* __restrict__ Xs = A.Xs;
  const float* __restrict__ Xf = A.Xf;
  const double* Gc = A.ws.G + (size_t)slot * p_pad;
  const double* XMc = A.ws.XM + (size_t)pid * p_pad;
  const double* XIc = A.ws.XI + (size_t)pid * p_pad;
  int* MMc = A.ws.MM + (size_t)pid * p_pad;
  int* SIc = A.ws.SIDX + (size_t)pid * p_pad;
  int* IAg = A.ws.IA + (size_t)pid * kNX;
  double* AAg = A.ws.AACT + (size_t)pid * kNX;
  double* as_ = A.ws.AS + (size_t)pid * kNX;  // coefficients at the start of the IRLS step (global)
  double* Wg = A.ws.W + (size_t)pid * ldx;
  double* WRg = A.ws.WR + (size_t)pid * ldx;
  const double* y = A.Yn + (size_t)g * n;
  const int* fo = A.fold + (size_t)g * n;
  const int self = A.self ? A.self[g] : -1;
  const int nvars = p - (self >= 0 ? 1 : 0);
  int nx = 2 * A.prm.dfmax + 20;
  if (nx > nvars) nx = nvars;
  if (nx > kNX) nx = kNX;
  const int ne = A.prm.dfmax;
  const int nlam = A.ws.nlam_g[g];
  const double* lamg = A.ws.lam + (size_t)g * kNLam;
  const bool master_auto = A.prm.auto_lambda && fit == 0;
  const bool is_fold = fit > 0;
  const int maxit = A.prm.maxit;
  const double fmax_ = log(DBL_MAX * 0.1);
  int crumb_m = 0;
  auto crumb = [&](int tag) {
    if (crumb_p && tid == 0) {
      volatile int* c = crumb_p;
      c[0] = A.round_no; c[1] = pid; c[2] = tag; c[3] = lidx; c[4] = nin; c[5] = nS; c[6] = crumb_m;
      c[7] = ninact;
      __threadfence_system();
    }
  };
  crumb(1);

  if (!STAGED) {
    wv = Wg;
    wrv = WRg;
  } else {
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      *reinterpret_cast<double2*>(wv + i) = __ldcs(reinterpret_cast<const double2*>(Wg + i));
      *reinterpret_cast<double2*>(wrv + i) = __ldcs(reinterpret_cast<const double2*>(WRg + i));
    }
  }
  for (int l = tid; l < nin; l += T) {
    const int k = IAg[l];
    ia[l] = k;
    aact[l] = AAg[l];
    xma[l] = XMc[k];
    xia[l] = XIc[k];
  }
  __syncthreads();

  // ---------------- phase A: consume the gradient of the previous round ----------------
  auto mark_pass = [&](double thresh) -> int {
    int cnt = 0;
    for (int j = tid; j < p; j += T) {
      const double xi = __ldcs(XIc + j);
      if (xi != 0.0 && MMc[j] == 0) {
        const double ga = fabs((__ldcs(Gc + j) - __ldcs(XMc + j) * sumr) * xi);
        if (ga > thresh) {
          MMc[j] = -1;
          ++cnt;
        }
      }
    }
    return __syncthreads_or(cnt);
  };

  // On a lambda advance the strong set is rebuilt from the sequential strong rule on the fresh
  // gradient (actives + {ga > 2 lambda_new - lambda_old}) instead of fishnet1's ever-growing ixx:
  // the full KKT check of the next round still covers every predictor, so the solution is the
  // same; the per-step strong-set check just stops paying for hundreds of long-dead candidates.
  auto mark_reset = [&](double thresh) -> int {
    int cnt = 0;
    for (int j = tid; j < p; j += T) {
      const double xi = __ldcs(XIc + j);
      const int mm = MMc[j];
      if (xi != 0.0 && mm <= 0) {
        const double ga = fabs((__ldcs(Gc + j) - __ldcs(XMc + j) * sumr) * xi);
        const int nm = ga > thresh ? -1 : 0;
        if (nm != mm) {
          MMc[j] = nm;
          ++cnt;
        }
      }
    }
    return __syncthreads_or(cnt);
  };
  bool strong_changed = false;
  if (state == ST_INIT && (nlam < 1 || A.ws.prob[(size_t)g * A.prm.NF].jerr > 0)) {
    state = ST_DONE;  // the gene's master fit failed: the whole gene falls back to the mean
  } else if (state == ST_INIT) {
    if (master_auto) {
      double m = 0.0;
      for (int j = tid; j < p; j += T) {
        const double xi = __ldcs(XIc + j);
        if (xi != 0.0) m = fmax(m, fabs((__ldcs(Gc + j) - __ldcs(XMc + j) * sumr) * xi));
      }
      m = bmax(m, red, phase);
      if (tid == 0) {  // the null model is solution 0 (lambda = "big")
        A.ws.a0p[(size_t)g * kNLam] = az;
        A.ws.devp[(size_t)g * kNLam] = cur_dev;
        A.ws.ninp[(size_t)g * kNLam] = 0;
      }
      lmu = 1;
      lidx = 1;
      al0 = m;
      al = lamg[1];
      if (nlam < 2) state = ST_DONE;
    } else {
      lidx = 0;
      al0 = 0.0;
      al = lamg[0];
    }
    if (state != ST_DONE) {
      strong_changed = mark_pass(2.0 * al - al0) != 0;
      state = ST_SOLVE;
    }
  } else if (state == ST_WAIT_KKT || state == ST_HOLD) {
    if (state == ST_WAIT_KKT && mark_pass(al)) {
      strong_changed = true;
      state = ST_SOLVE;  // KKT violated outside the strong set: re-solve this lambda
    } else {
      bool stop = false;
      if (state == ST_WAIT_KKT) {
      // ---- commit solution lidx ----
      if (fit == 0) {
        const size_t o = (size_t)g * kNLam + lidx;
        if (tid == 0) {
          A.ws.a0p[o] = az;
          A.ws.devp[o] = cur_dev;
          A.ws.ninp[o] = nin;
          A.ws.nlpp[o] = nlp;
        }
        double* cp = A.ws.cap + o * kNX;
        for (int l = tid; l < nin; l += T) cp[l] = aact[l];
      } else if (tid == 0) {
        A.ws.cvdev[((size_t)g * kNLam + lidx) * kMaxFolds + (fit - 1)] = cur_cv / ncv;
      }
      lmu = lidx + 1;
      if (master_auto) {
        const int mnl = kMnLam < nlam ? kMnLam : nlam;
        if (lidx + 1 >= mnl) {
          int me = 0;
          for (int l = tid; l < nin; l += T) me += aact[l] != 0.0;
          double mv[1] = {(double)me};
          bsum<1>(mv, red, phase);
          const double dprev = A.ws.devp[(size_t)g * kNLam + lidx - mnl + 1];
          if (mv[0] > ne) stop = true;
          else if ((cur_dev - dprev) / cur_dev < kSml) stop = true;
          else if (cur_dev > kDevMax) stop = true;
        }
      }
      }
      int lim = nlam;
      bool hold = false;
      if (is_fold && A.prm.auto_lambda && !(A.prm.dbg & 8192)) {  // bit 8192: folds ignore the master (fault hunting)
        const int mlmu = A.ws.msnap[2 * g], mdone = A.ws.msnap[2 * g + 1];
        if (mdone >= 0) { if (mdone < lim) lim = mdone; }
        else if (lidx + 1 > mlmu + 1) hold = true;  // at most one lambda ahead of the master's path
      }
      if (stop || lidx + 1 >= lim) {
        state = ST_DONE;
      } else if (hold) {
        state = ST_HOLD;
      } else {
        ++lidx;
        al0 = al;
        al = lamg[lidx];
        strong_changed = mark_reset(al - A.prm.strong_slack * (al0 - al)) != 0;
        state = ST_SOLVE;
      }
    }
  }
  if (strong_changed) nS = rebuild_strong<T>(MMc, SIc, p, ish);

  // ---------------- phase B (Gram space): IRLS + coordinate descent on the strong set ----------
  // Within one IRLS step the working weights are fixed, so the cyclic sweeps over the active set
  // only need the weighted Gram block H = X~_A^T W X~_A and the gradient q = X~_A^T wr: a
  // coordinate update is q -= d * H[:, l] on an |A|-vector instead of an n-long dot + axpy with a
  // block reduction.  H is built once per IRLS step with FP64 DMMA straight from the resident
  // design; the n-space residual is only materialised before the strong-set check of the inactive
  // variables (batched dots, no per-variable barrier) and when a variable enters.
  lap(0);
  crumb(2);
  double* H = nullptr;
  int hslot = -1;
  if (state == ST_SOLVE) {
    if (tid == 0) {
      int sidx = blockIdx.x % kHSlots;
      while (atomicCAS(A.ws.hflags + sidx, 0, 1) != 0) sidx = (sidx + 1) % kHSlots;
      ish[41] = sidx;
    }
    __syncthreads();
    hslot = ish[41];
    H = A.ws.Hs + (size_t)hslot * kNX * kNX;
  }
  constexpr int ldh = kNX;
  constexpr int EPT = (kNX + T - 1) / T;
  double q0 = 0.0, azmat = az;
  double v0m = v0;       // sum of the weights the Gram block was built with (the model's v0)
  bool H_valid = false;  // a Gram block exists for this launch
  bool H_fresh = false;  // ... and it was built with the current working weights
  bool dirty = false;

  // A streaming phase over ng groups of 4 columns costs one column pass per round of T/32 groups,
  // however few groups there are (a single warp cannot pull more than ~2 GB/s; the SM reaches its
  // share of HBM only with every warp streaming: profiles/r2_ubench_stream.log).  Short phases are
  // therefore cut into (group, row segment) items: S segments per column, chosen to minimise
  // rounds x segment length, partial sums combined through `part`.
  auto pick_split = [&](int ng, int maxitems) -> int {
    if (A.prm.dbg & 65536) return 1;
    const int nw = T >> 5;
    int best = 1, bc = (ng + nw - 1) / nw * 16;
    for (int s2 = 2; s2 <= 16; s2 *= 2) {
      if (ng * s2 > maxitems || s2 * 512 > ldx) break;
      const int c = (ng * s2 + nw - 1) / nw * (16 / s2);
      if (c < bc) { bc = c; best = s2; }
    }
    return best;
  };
  auto seg_len = [&](int S) { return ((ldx + S - 1) / S + 127) & ~127; };
  // p_l = sum w Xs_l -> swa[l], r_l = sum wr Xs_l -> gq[l]   (four columns per reduction)
  // (one warp per column group / row segment, independent accumulators per lane)
  auto gram_vectors = [&](int l0, int l1, bool with_p) {
    crumb_m = l1;
    crumb(17);
    const int lane = tid & 31, warp = tid >> 5, nw = T >> 5;
    if (mlp) {
      const int ng = (l1 - l0 + 3) >> 2, S = pick_split(ng, kPart / 8), sl = seg_len(S);
      for (int item = warp; item < ng * S; item += nw) {
        const int grp = item / S, seg = item - grp * S, lb = l0 + 4 * grp;
        const int r0 = seg * sl, r1 = r0 + sl < ldx ? r0 + sl : ldx;
        const double* zc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) zc[c] = Xs + (size_t)ia[lb + c < l1 ? lb + c : lb] * ldx;
        double ps4[4], rs4[4];
        wdot_wr4(zc[0], zc[1], zc[2], zc[3], wv, wrv, r0, r1, ps4, rs4);
        if (lane == 0) {
          if (S == 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (lb + c < l1) {
                if (with_p) swa[lb + c] = ps4[c];
                gq[lb + c] = rs4[c];
              }
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              part[item * 8 + c] = ps4[c];
              part[item * 8 + 4 + c] = rs4[c];
            }
          }
        }
      }
      if (S > 1) {
        __syncthreads();
        for (int l = l0 + tid; l < l1; l += T) {
          const int grp = (l - l0) >> 2, c = (l - l0) & 3;
          double ps = 0.0, rs = 0.0;
          for (int seg = 0; seg < S; ++seg) {
            ps += part[(grp * S + seg) * 8 + c];
            rs += part[(grp * S + seg) * 8 + 4 + c];
          }
          if (with_p) swa[l] = ps;
          gq[l] = rs;
        }
      }
    } else
    for (int l = l0 + warp; l < l1; l += nw) {
      const double* c = Xs + (size_t)ia[l] * ldx;
      double pa = 0.0, pb = 0.0, ra = 0.0, rb = 0.0;
#pragma unroll 4
      for (int i = 2 * lane; i < ldx; i += 64) {
        const double2 x = ldg2(c + i);
        const double2 w = *reinterpret_cast<const double2*>(wv + i);
        const double2 r = *reinterpret_cast<const double2*>(wrv + i);
        pa += w.x * x.x;
        pb += w.y * x.y;
        ra += r.x * x.x;
        rb += r.y * x.y;
      }
      const double ps = warp_sum(pa + pb), rs = warp_sum(ra + rb);
      if (lane == 0) {
        if (with_p) swa[l] = ps;
        gq[l] = rs;
      }
    }
    ncols += l1 - l0;
    __syncthreads();
  };
  // raw weighted Gram P[l][k] = sum_i w_i Xs_il Xs_ik for l, k < m, 16 x 16 tiles per warp on DMMA
  auto gram_P = [&](int m) {
    crumb_m = m;
    crumb(16);
    const int lane = tid & 31, warp = tid >> 5, nw = T >> 5, g8 = lane >> 2, t4 = lane & 3;
    const int T16 = (m + 15) >> 4, ntile = T16 * (T16 + 1) / 2;
    // ntile tiles on nw warps leave the CTA waiting for the warps with one tile more (16 % of all
    // stall samples sat at this phase's closing barrier, profiles/r2_ncu_lasso_lines.txt): tiles are
    // cut into S row segments, S minimising rounds / S, partial tiles summed in a fixed order.
    int S = 1;
    if (!(A.prm.dbg & 65536)) {
      int bn = (ntile + nw - 1) / nw;  // best cost = bn / S
      for (int s2 = 2; s2 <= 16; ++s2) {
        if (ntile * s2 > kGramItems || s2 * 512 > ldx) break;
        const int c = (ntile * s2 + nw - 1) / nw;
        if (c * S < bn * s2) { bn = c; S = s2; }
      }
    }
    const int sl = ((ldx + S - 1) / S + 15) & ~15;
    double* HPc = A.ws.HP + (size_t)hslot * kGramItems * 256;
    auto decode = [&](int tix, int& I, int& J) {
      J = (int)((sqrt(8.0 * tix + 1.0) - 1.0) * 0.5);
      while ((J + 1) * (J + 2) / 2 <= tix) ++J;
      while (J * (J + 1) / 2 > tix) --J;
      I = tix - J * (J + 1) / 2;
    };
    for (int item = warp; item < ntile * S; item += nw) {
      const int tix = item / S, seg = item - tix * S;
      const int r0 = seg * sl, r1 = r0 + sl < ldx ? r0 + sl : ldx;
      int I, J;
      decode(tix, I, J);
      double acc[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
      if (Xf) {  // FP32 copy: 16 rows per step, four k-slices of one 16-byte load per lane
        auto colf = [&](int idx) { return Xf + (size_t)ia[idx < m ? idx : m - 1] * ldx + 4 * t4; };
        const float *fI0 = colf(I * 16 + g8), *fI1 = colf(I * 16 + 8 + g8);
        const float *fJ0 = colf(J * 16 + g8), *fJ1 = colf(J * 16 + 8 + g8);
#pragma unroll 2
        for (int base = r0; base < r1; base += 16) {
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(fI0 + base));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(fI1 + base));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(fJ0 + base));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(fJ1 + base));
          const double2 wa = *reinterpret_cast<const double2*>(wv + base + 4 * t4);
          const double2 wb = *reinterpret_cast<const double2*>(wv + base + 4 * t4 + 2);
          const double av0[4] = {(double)a0.x, (double)a0.y, (double)a0.z, (double)a0.w};
          const double av1[4] = {(double)a1.x, (double)a1.y, (double)a1.z, (double)a1.w};
          const double bv0[4] = {wa.x * (double)b0.x, wa.y * (double)b0.y, wb.x * (double)b0.z, wb.y * (double)b0.w};
          const double bv1[4] = {wa.x * (double)b1.x, wa.y * (double)b1.y, wb.x * (double)b1.z, wb.y * (double)b1.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dmma(acc[0][0][0], acc[0][0][1], av0[j], bv0[j]);
            dmma(acc[0][1][0], acc[0][1][1], av0[j], bv1[j]);
            dmma(acc[1][0][0], acc[1][0][1], av1[j], bv0[j]);
            dmma(acc[1][1][0], acc[1][1][1], av1[j], bv1[j]);
          }
        }
      } else {
        auto colp = [&](int idx) { return Xs + (size_t)ia[idx < m ? idx : m - 1] * ldx + 2 * t4; };
        const double *cI0 = colp(I * 16 + g8), *cI1 = colp(I * 16 + 8 + g8);
        const double *cJ0 = colp(J * 16 + g8), *cJ1 = colp(J * 16 + 8 + g8);
#pragma unroll 2
        for (int base = r0; base < r1; base += 8) {
          const double2 a0 = ldg2(cI0 + base), a1 = ldg2(cI1 + base);
          double2 b0 = ldg2(cJ0 + base), b1 = ldg2(cJ1 + base);
          const double2 w = *reinterpret_cast<const double2*>(wv + base + 2 * t4);
          b0.x *= w.x; b0.y *= w.y; b1.x *= w.x; b1.y *= w.y;
          dmma(acc[0][0][0], acc[0][0][1], a0.x, b0.x);
          dmma(acc[0][1][0], acc[0][1][1], a0.x, b1.x);
          dmma(acc[1][0][0], acc[1][0][1], a1.x, b0.x);
          dmma(acc[1][1][0], acc[1][1][1], a1.x, b1.x);
          dmma(acc[0][0][0], acc[0][0][1], a0.y, b0.y);
          dmma(acc[0][1][0], acc[0][1][1], a0.y, b1.y);
          dmma(acc[1][0][0], acc[1][0][1], a1.y, b0.y);
          dmma(acc[1][1][0], acc[1][1][1], a1.y, b1.y);
        }
      }
      if (S == 1) {
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int r = I * 16 + a * 8 + g8, cc = J * 16 + b * 8 + 2 * t4 + e;
              if (r < m && cc < m && (I != J || r <= cc)) {
                H[r + (size_t)cc * ldh] = acc[a][b][e];
                H[cc + (size_t)r * ldh] = acc[a][b][e];
              }
            }
      } else {
        double2* dst = reinterpret_cast<double2*>(HPc + ((size_t)item * 32 + lane) * 8);
        dst[0] = make_double2(acc[0][0][0], acc[0][0][1]);
        dst[1] = make_double2(acc[0][1][0], acc[0][1][1]);
        dst[2] = make_double2(acc[1][0][0], acc[1][0][1]);
        dst[3] = make_double2(acc[1][1][0], acc[1][1][1]);
      }
    }
    gram_fma += (unsigned long long)ntile * 256ull * (unsigned)ldx;
    __syncthreads();
    if (S > 1) {
      for (int idx = tid; idx < ntile * 256; idx += T) {
        const int tix = idx >> 8, ln = (idx >> 3) & 31, k = idx & 7;
        double v = 0.0;
        for (int seg = 0; seg < S; ++seg) v += HPc[((size_t)(tix * S + seg) * 32 + ln) * 8 + k];
        int I, J;
        decode(tix, I, J);
        const int r = I * 16 + (k >> 2) * 8 + (ln >> 2), cc = J * 16 + ((k >> 1) & 1) * 8 + 2 * (ln & 3) + (k & 1);
        if (r < m && cc < m && (I != J || r <= cc)) {
          H[r + (size_t)cc * ldh] = v;
          H[cc + (size_t)r * ldh] = v;
        }
      }
      __syncthreads();
    }
  };
  // per-fit standardisation of the raw moments: H and v (-> va) ...
  auto gram_finish_H = [&](int m) {
    const bool pack = m <= mh;
    for (int idx = tid; idx < m * m; idx += T) {
      const int r = idx % m, c = idx / m;
      const double P = H[r + (size_t)c * ldh];
      const double hv =
          xia[r] * xia[c] * (P - xma[r] * swa[c] - xma[c] * swa[r] + xma[r] * xma[c] * v0);
      H[r + (size_t)c * ldh] = hv;
      if (pack && r >= c) Hsm[r * (r + 1) / 2 + c] = hv;
    }
    hs_m = pack ? m : -1;
    __syncthreads();
    for (int l = tid; l < m; l += T) va[l] = H[l + (size_t)l * ldh];
    __syncthreads();
  };
  // ... and of the vectors: q (-> gq) always, h0 (-> swa) when the block was rebuilt
  auto gram_finish_vec = [&](int m, bool with_p) {
    for (int l = tid; l < m; l += T) {
      gq[l] = xia[l] * (gq[l] - xma[l] * q0);  // q0 = current sum of the working residual
      if (with_p) swa[l] = xia[l] * (swa[l] - xma[l] * v0);
    }
    __syncthreads();
  };
  // One cyclic sweep over the active set in Gram space.  For m <= 256 the whole sweep runs inside
  // warp 0 with q in registers (8 entries per lane): a coordinate step is a shuffle broadcast, a
  // few dependent FP64 operations and 8 FMAs per lane, with the Hessian column prefetched two
  // steps ahead -- no block barrier inside the sweep.  Larger active sets use the block version.
  auto gsweep_warp = [&](bool by_index, int m, double& dlx) {
    if (tid < 32) {
      if (hs_m == m) {
        if (m <= 32) gram_sweep_sm<1>(Hsm, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
        else if (m <= 64) gram_sweep_sm<2>(Hsm, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
        else if (m <= 96) gram_sweep_sm<3>(Hsm, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
        else if (m <= 128) gram_sweep_sm<4>(Hsm, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
        else if (m <= 192) gram_sweep_sm<6>(Hsm, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
        else gram_sweep_sm<8>(Hsm, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
      } else if (m <= 64) gram_sweep_g<2>(H, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
      else if (m <= 128) gram_sweep_g<4>(H, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
      else if (m <= 192) gram_sweep_g<6>(H, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
      else if (m <= 256) gram_sweep_g<8>(H, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
      else if (m <= 320) gram_sweep_g<10>(H, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
      else gram_sweep_g<12>(H, m, by_index, ord, gq, va, gq2, aact, swa, al, gsh, dlx, q0);
    }
    __syncthreads();
    dlx = gsh[0];
    q0 = gsh[1];
    if (gsh[2] != 0.0) dirty = true;
    __syncthreads();
  };
  auto gsweep_block = [&](bool by_index, int m, double& dlx) {
    double hreg[EPT];
    int lnext = by_index ? ord[0] : 0;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int j = tid + e * T;
      hreg[e] = j < m ? H[j + (size_t)lnext * ldh] : 0.0;
    }
    // q is double buffered (gq / gq2): a changed step reads the current buffer and writes the
    // other one, so one barrier per changed step orders everything
    double* qa = gq;
    double* qb = gq2;
    for (int t = 0; t < m; ++t) {
      const int l = lnext;
      double hcur[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) hcur[e] = hreg[e];
      if (t + 1 < m) {
        lnext = by_index ? ord[t + 1] : t + 1;
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          const int j = tid + e * T;
          hreg[e] = j < m ? H[j + (size_t)lnext * ldh] : 0.0;
        }
      }
      const double ql = qa[l], v = va[l], ak = aact[l];
      const double u = ql + v * ak;
      const double au = fabs(u) - al;
      const double anew = au > 0.0 ? copysign(au, u) / v : 0.0;
      if (anew == ak) continue;
      if (anew != anew) { dlx = NAN; break; }
      const double d = anew - ak;
      dlx = fmax(dlx, v * d * d);
      q0 -= d * swa[l];
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int j = tid + e * T;
        if (j < m) qb[j] = qa[j] - d * hcur[e];
      }
      dirty = true;
      __syncthreads();
      if (tid == 0) aact[l] = anew;  // nobody reads aact[l] again before the next barrier
      double* tq = qa; qa = qb; qb = tq;
    }
    __syncthreads();
    if (qa != gq) {  // leave the result in gq
      for (int j = tid; j < m; j += T) gq[j] = qa[j];
    }
    __syncthreads();
  };
  auto gsweep = [&](bool by_index, int m, double& dlx) {
    if (m == 0) return;
    tph[8] += m;
    tph[9] += 1;
    if (timing && tid == 0) atomicAdd(A.ws.phasecyc + 16 + (m - 1) / 32, (unsigned long long)m);
    crumb_m = m;
    crumb(m <= 384 ? (hs_m == m ? 11 : 12) : 13);
    if (m <= 384) gsweep_warp(by_index, m, dlx);
    else gsweep_block(by_index, m, dlx);
    crumb(14);
  };
  auto gintercept = [&](int m, double& dlx) {
    const double d = q0 / v0m;
    az += d;
    dlx = fmax(dlx, v0m * d * d);
    for (int j = tid; j < m; j += T) gq[j] -= d * swa[j];
    q0 -= d * v0m;
    if (d != 0.0) dirty = true;
    __syncthreads();
  };
  // Order-preserving compaction of the non-zero entries of cs[0..m) (in place) and of their design
  // columns (-> cks, aliasing `part`): the row combinations below skip members whose coefficient
  // is (or stayed) zero -- exact zeros dropped from a sum leave it bit-identical.
  int* cks = reinterpret_cast<int*>(part);
  auto compact_nonzero = [&](double* cs, int m) -> int {
    __syncthreads();
    if (tid < 32) {
      int base = 0;
      for (int l0 = 0; l0 < m; l0 += 32) {
        const int l = l0 + tid;
        const double c = l < m ? cs[l] : 0.0;
        const int k = l < m ? ia[l] : 0;
        const unsigned b = __ballot_sync(0xffffffffu, c != 0.0);
        const int pos = base + __popc(b & ((1u << tid) - 1u));
        __syncwarp();
        if (c != 0.0) { cs[pos] = c; cks[pos] = k; }
        base += __popc(b);
      }
      if (tid == 0) ish[43] = base;
    }
    __syncthreads();
    return ish[43];
  };
  // bring the n-space residual up to date with the coefficients: wr -= w * (X~_A da + daz)
  auto materialise = [&](int m) {
    if (!dirty) return;
    crumb_m = m;
    crumb(18);
    double sh[1] = {0.0};
    for (int l = tid; l < m; l += T) {
      const double c = (aact[l] - amat[l]) * xia[l];
      amat[l] = c;
      sh[0] += c * xma[l];
    }
    bsum<1>(sh, red, phase);
    const double shift = (az - azmat) - sh[0];
    const int mc = mlp ? compact_nonzero(amat, m) : m;  // members whose coefficient did not move add nothing
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      double s0 = shift, s1 = shift;
      if (mlp) {
        const double2 sc = rows_comb(Xs + i, ldx, cks, amat, mc, s0, s1);
        s0 = sc.x;
        s1 = sc.y;
      } else
#pragma unroll 4
      for (int l = 0; l < m; ++l) {
        const double c = amat[l];
        if (c != 0.0) {
          const double2 x = ldg2(Xs + (size_t)ia[l] * ldx + i);
          s0 += c * x.x;
          s1 += c * x.y;
        }
      }
      const double2 w = *reinterpret_cast<const double2*>(wv + i);
      double2 r = *reinterpret_cast<const double2*>(wrv + i);
      r.x -= w.x * s0;
      r.y -= w.y * s1;
      *reinterpret_cast<double2*>(wrv + i) = r;
    }
    ncols += mc;
    __syncthreads();
    for (int l = tid; l < m; l += T) amat[l] = aact[l];
    azmat = az;
    dirty = false;
    __syncthreads();
  };
  // a strong-set variable enters: n-space coordinate update + one new Gram row/column
  auto activate = [&](int k, double xm, double xi, double gk, int& m, double& dlx) -> bool {
    crumb_m = m;
    crumb(7);
    const double* col = Xs + (size_t)k * ldx;
    double d2[2] = {0.0, 0.0};
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      const double2 x = ldg2(col + i);
      const double2 w = *reinterpret_cast<const double2*>(wv + i);
      const double wx0 = w.x * x.x, wx1 = w.y * x.y;
      d2[0] += wx0 + wx1;
      d2[1] += wx0 * x.x + wx1 * x.y;
    }
    bsum<2>(d2, red, phase);
    const double sw = xi * (d2[0] - xm * v0);
    const double v = xi * xi * (d2[1] - 2.0 * xm * d2[0] + xm * xm * v0);
    const double au = fabs(gk) - al;
    const double d = copysign(au, gk) / v;
    if (d != d) { dlx = NAN; return false; }
    dlx = fmax(dlx, v * d * d);
    {  // wr -= d * w * x~_k
      const double c = d * xi;
      for (int i = 2 * tid; i < ldx; i += 2 * T) {
        const double2 x = ldg2(col + i);
        double2 r = *reinterpret_cast<const double2*>(wrv + i);
        const double2 w = *reinterpret_cast<const double2*>(wv + i);
        r.x -= c * w.x * (x.x - xm);
        r.y -= c * w.y * (x.y - xm);
        *reinterpret_cast<double2*>(wrv + i) = r;
      }
    }
    q0 -= d * sw;
    ++nin;
    if (nin > nx) return false;  // pmax overflow, handled by the caller
    const bool grow = hs_m == m && m + 1 <= mh;  // the packed block takes row m as well
    double* hrow = Hsm + m * (m + 1) / 2;
    // new Gram column against the m current members: one warp per member, no block barrier
    {
      const int lane = tid & 31, warp = tid >> 5, nw = T >> 5;
      auto put = [&](int l, double acc) {  // lane 0: standardise, store the new row / column
        const double pl = swa[l] / xia[l] + xma[l] * v0m;  // raw sum w Xs_l back from h0_l
        const double h = xia[l] * xi * (acc - xma[l] * d2[0] - xm * pl + xma[l] * xm * v0);
        H[l + (size_t)m * ldh] = h;
        H[m + (size_t)l * ldh] = h;
        if (grow) hrow[l] = h;
        gq[l] -= d * h;
      };
      if (mlp && Xf) {
        const int ng = (m + 3) >> 2, S = pick_split(ng, kPart / 4), sl = seg_len(S);
        for (int item = warp; item < ng * S; item += nw) {
          const int grp = item / S, seg = item - grp * S, lb = 4 * grp;
          const int r0 = seg * sl, r1 = r0 + sl < ldx ? r0 + sl : ldx;
          const float* zf[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) zf[c] = Xf + (size_t)ia[lb + c < m ? lb + c : lb] * ldx;
          double acc4[4];
          wdot_wxz4f(zf[0], zf[1], zf[2], zf[3], Xf + (size_t)k * ldx, wv, r0, r1, acc4);
          if (lane == 0) {
            if (S == 1) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (lb + c < m) put(lb + c, acc4[c]);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) part[item * 4 + c] = acc4[c];
            }
          }
        }
        if (S > 1) {
          __syncthreads();
          for (int l = tid; l < m; l += T) {
            double acc = 0.0;
            for (int seg = 0; seg < S; ++seg) acc += part[((l >> 2) * S + seg) * 4 + (l & 3)];
            put(l, acc);
          }
        }
      } else if (mlp) {
        for (int lb = 4 * warp; lb < m; lb += 4 * nw) {
          double acc4[4];
          {
          const double* zc[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) zc[c] = Xs + (size_t)ia[lb + c < m ? lb + c : lb] * ldx;
          wdot_wxz4(zc[0], zc[1], zc[2], zc[3], col, wv, ldx, acc4);
          }
          if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (lb + c < m) put(lb + c, acc4[c]);
          }
        }
      } else {
        for (int l = warp; l < m; l += nw) {
          const double* c = Xs + (size_t)ia[l] * ldx;
          double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
          for (int i = 2 * lane; i < ldx; i += 64) {
            const double2 x = ldg2(col + i);
            const double2 z = ldg2(c + i);
            const double2 w = *reinterpret_cast<const double2*>(wv + i);
            s0 += w.x * x.x * z.x;
            s1 += w.y * x.y * z.y;
          }
          const double acc = warp_sum(s0 + s1);
          if (lane == 0) put(l, acc);
        }
      }
      ncols += m;
      __syncthreads();
    }
    hs_m = grow ? m + 1 : -1;
    if (tid == 0) {
      H[m + (size_t)m * ldh] = v;
      if (grow) hrow[m] = v;
      MMc[k] = nin;
      ia[m] = k;
      xma[m] = xm;
      xia[m] = xi;
      va[m] = v;
      gq2[m] = 1.0 / v;
      swa[m] = sw;
      as_[m] = 0.0;
      aact[m] = d;
      amat[m] = d;
      gq[m] = gk - d * v;
      int t = m;  // keep ord[] sorted by variable index
      while (t > 0 && ia[ord[t - 1]] > k) { ord[t] = ord[t - 1]; --t; }
      ord[t] = m;
    }
    ncols += 2;
    m = nin;
    __syncthreads();
    return true;
  };
  // Strong-set check of the inactive variables (KKT inside the strong set): g_k = x~_k^T wr for a
  // staged chunk of candidates, one warp per column; the first variable (in index order) whose
  // |g_k| exceeds lambda enters, after which the rest of the chunk is evaluated again.
  const bool screen = Xf != nullptr && !(A.prm.dbg & 32768);  // bit 32768: FP64 strong-set dots only
  auto inactive_check = [&](int& m, double& dlx, bool& overflow) {
    if (ninact == 0) return;
    tph[11] += 1;
    materialise(m);
    lap(4);
    const int lane = tid & 31, warp = tid >> 5, nw = T >> 5;
    int cur = 0;
    while (cur < ninact) {
      const int cn = ninact - cur < kMeta ? ninact - cur : kMeta;
      __syncthreads();
      if (tid < cn) {
        const int k = INc[cur + tid];
        mk[tid] = k;
        mxm[tid] = XMc[k];
        mxi[tid] = XIc[k];
        mpos[tid] = MMc[k];
      }
      __syncthreads();
      int j0 = 0;
      crumb_m = cur;
      crumb(15);
      while (j0 < cn) {
        if (mlp && screen) {
          // screen on the FP32 copy: |g| is decided against lambda by an interval that contains the
          // FP64 value; the undecided and the violators get the FP64 dot below
          const int ng = (cn - j0 + 3) >> 2, S = pick_split(ng, kPart / 8), sl = seg_len(S);
          for (int item = warp; item < ng * S; item += nw) {
            const int grp = item / S, seg = item - grp * S, jb = j0 + 4 * grp;
            const int r0 = seg * sl, r1 = r0 + sl < ldx ? r0 + sl : ldx;
            bool need[4], any = false;
            int first = -1;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              need[c] = jb + c < cn && mpos[jb + c] < 0;
              any |= need[c];
              if (need[c] && first < 0) first = jb + c;
            }
            if (!any) continue;
            const float* zf[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) zf[c] = Xf + (size_t)mk[need[c] ? jb + c : first] * ldx;
            double acc4[4], abs4[4];
            wdot_a4f(zf[0], zf[1], zf[2], zf[3], wrv, r0, r1, acc4, abs4);
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                part[item * 8 + c] = acc4[c];
                part[item * 8 + 4 + c] = abs4[c];
              }
            }
          }
          __syncthreads();
          for (int j = j0 + tid; j < cn; j += T)
            if (mpos[j] < 0) {
              const int grp = (j - j0) >> 2, c = (j - j0) & 3;
              double acc = 0.0, ab = 0.0;
              for (int seg = 0; seg < S; ++seg) {
                acc += part[(grp * S + seg) * 8 + c];
                ab += part[(grp * S + seg) * 8 + 4 + c];
              }
              const double gf = mxi[j] * (acc - mxm[j] * q0);
              const double bnd = mxi[j] * ab * 6.0e-8 + fabs(gf) * 1e-13;
              mg[j] = gf;
              mamb[j] = !(fabs(gf) + bnd <= al);
            }
          __syncthreads();
          // FP64 dots of the flagged candidates, every warp on its row segment of one column
          int S2 = 1;
          while (S2 < nw && 2 * S2 * 512 <= ldx) S2 *= 2;
          const int sl2 = seg_len(S2);
          for (int j = j0; j < cn; ++j)
            if (mpos[j] < 0 && mamb[j]) {
              if (warp < S2) {
                const int r0 = warp * sl2, r1 = r0 + sl2 < ldx ? r0 + sl2 : ldx;
                const double acc = wdot_a(Xs + (size_t)mk[j] * ldx, wrv, r0, r1);
                if (lane == 0) part[warp] = acc;
              }
              __syncthreads();
              if (tid == 0) {
                double acc = 0.0;
                for (int seg = 0; seg < S2; ++seg) acc += part[seg];
                mg[j] = mxi[j] * (acc - mxm[j] * q0);
              }
              tph[15] += 1;
              __syncthreads();
              // the scan below stops at the first violator in index order: later candidates are
              // evaluated again after it has entered, their FP64 value is not needed now
              if (fabs(mg[j]) > al) break;
            }
        } else if (mlp) {
          for (int jb = j0 + 4 * warp; jb < cn; jb += 4 * nw) {
            const double* zc[4];
            bool need[4], any = false;
            int first = -1;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              need[c] = jb + c < cn && mpos[jb + c] < 0;
              any |= need[c];
              if (need[c] && first < 0) first = jb + c;
            }
            if (!any) continue;
            double acc4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) zc[c] = Xs + (size_t)mk[need[c] ? jb + c : first] * ldx;
            wdot_a4(zc[0], zc[1], zc[2], zc[3], wrv, ldx, acc4);
            if (lane == 0) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (need[c]) mg[jb + c] = mxi[jb + c] * (acc4[c] - mxm[jb + c] * q0);
            }
          }
        } else {
          for (int j = j0 + warp; j < cn; j += nw) {
            if (mpos[j] < 0) {
              const double* c = Xs + (size_t)mk[j] * ldx;
              double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
              for (int i = 2 * lane; i < ldx; i += 64) {
                const double2 x = ldg2(c + i);
                const double2 r = *reinterpret_cast<const double2*>(wrv + i);
                s0 += r.x * x.x;
                s1 += r.y * x.y;
              }
              const double acc = warp_sum(s0 + s1);
              if (lane == 0) mg[j] = mxi[j] * (acc - mxm[j] * q0);
            }
          }
        }
        __syncthreads();
        int hit = -1;
        for (int j = j0; j < cn; ++j)
          if (mpos[j] < 0) {
            ++ncols;
            tph[10] += 1;
            if (fabs(mg[j]) > al) { hit = j; break; }
          }
        if (hit < 0) break;
        lap(5);
        tph[14] += 1;
        if (!activate(mk[hit], mxm[hit], mxi[hit], mg[hit], m, dlx)) {
          if (nin > nx) overflow = true;
          return;
        }
        lap(13);
        if (tid == 0) mpos[hit] = nin;
        __syncthreads();
        j0 = hit + 1;  // the rest of the chunk sees the updated residual
      }
      cur += cn;
    }
  };

  if (state == ST_SOLVE) {
    // inactive members of the strong set, in index order
    {
      int base = 0;
      const int lane = tid & 31, warp = tid >> 5, nw = T >> 5;
      for (int s0 = 0; s0 < nS; s0 += T) {
        const int s = s0 + tid;
        const int k = s < nS ? SIc[s] : 0;
        const bool flag = s < nS && MMc[k] < 0;
        const unsigned b = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) ish[warp] = __popc(b);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < nw; ++w) {
          const int c = ish[w];
          if (w < warp) woff += c;
          tot += c;
        }
        if (flag) INc[base + woff + __popc(b & ((1u << lane) - 1u))] = k;
        base += tot;
        __syncthreads();
      }
      ninact = base;
    }
    // ord[] = active positions sorted by variable index
    for (int l = tid; l < nin; l += T) {
      int rank = 0;
      const int k = ia[l];
      for (int j = 0; j < nin; ++j) rank += ia[j] < k;
      ord[rank] = l;
    }
    __syncthreads();
    int irls_it = 0;
    bool big_step = false;
    while (true) {  // IRLS iterations
      const double az0 = az;
      for (int l = tid; l < nin; l += T) {
        as_[l] = aact[l];
        amat[l] = aact[l];
      }
      azmat = az;
      dirty = false;
      q0 = sumr;
      bool overflow = false, maxed = false, nanfail = false;
      int m = nin;
      __syncthreads();
      const int dbg = A.prm.dbg;
      if (m > 0 && !(dbg & 1)) {
        // The gradient is exact at every IRLS step.  The Hessian model (Gram block, intercept row
        // h0, v0) is built once per launch (= once per lambda) and kept for the later IRLS steps
        // of the same lambda as ONE consistent matrix (a true weighted Gram, so the sweeps cannot
        // cycle): the fixed point only depends on the gradient, the model only shapes the step
        // (chord iteration).  A variable entering under a stale model triggers a rebuild.
        lap(3);
        // ... unless the previous IRLS step was large (big_step) or the iteration drags on: with a
        // coarse lambda step (the fixed-lambda variant's factor e^-0.2) the weights move so far
        // from those of the kept model that the chord iteration oscillates up to maxit
        // (scripts/probe_heavy.py); then the model is rebuilt (plain IRLS, as fishnet1 does).
        const bool rebuild = !H_valid || big_step || irls_it >= kChordKeep || (dbg & 64);
        ++irls_it;
        gram_vectors(0, m, rebuild);
        lap(1);
        if (rebuild) {
          gram_P(m);
          gram_finish_H(m);
          v0m = v0;
          H_valid = true;
        }
        gram_finish_vec(m, rebuild);
        H_fresh = rebuild;
        for (int l = tid; l < m; l += T) gq2[l] = 1.0 / va[l];  // reciprocal curvatures for the warp sweep
        __syncthreads();
        lap(2);
      } else if (m == 0) {
        v0m = v0;
        H_fresh = true;
      }
      tph[7] += 1;
      tph[12] += m;
      crumb_m = m;
      crumb(4);
      while (true) {  // strong-set phases until one changes nothing
        ++nlp;
        double dlx = 0.0;
        if (!(dbg & 2)) gsweep(true, m, dlx);
        lap(3);
        lap(4);
        crumb_m = m;
        crumb(6);
        // The inactive candidates are looked at once the active set has settled: while the
        // index-order sweep still moves coefficients (dlx >= shr_in) the check -- a residual
        // materialisation plus one pass over every inactive strong column -- would judge an
        // intermediate residual; the phase cannot end without a check (dlx < shr_in is tested after
        // it), so the strong-set KKT conditions hold at exit as before.  dbg bit 131072: check in
        // every phase (fishnet1's order).
        const bool settled = !(dlx >= shr_in) || m == 0 || (dbg & 131072);
        if (!(dbg & 8) && settled) inactive_check(m, dlx, overflow);
        crumb_m = m;
        crumb(5);
        lap(5);
        if (overflow) break;
        gintercept(m, dlx);
        if (dlx < shr_in) break;
        if (!(dlx == dlx)) { nanfail = true; break; }
        if (nlp > maxit) { maxed = true; break; }
        int inner = 0;
        while (true) {  // sweeps over the ever-active list
          ++nlp;
          dlx = 0.0;
          if (!(dbg & 2)) gsweep(false, m, dlx);
          gintercept(m, dlx);
          if (dlx < shr_in) break;
          if (!(dlx == dlx)) { nanfail = true; break; }
          if (nlp > maxit) { maxed = true; break; }
          if (++inner >= 64 && !H_fresh && m > 0) {
            // A model that mixes weights of several IRLS steps (kept block + rows of variables that
            // entered later) is not guaranteed to be a Gram matrix; if the sweeps stall, rebuild it
            // with the current weights (then it is one, and coordinate descent converges).
            materialise(m);
            gram_vectors(0, m, true);
            gram_P(m);
            gram_finish_H(m);
            v0m = v0;
            gram_finish_vec(m, true);
            for (int l = tid; l < m; l += T) gq2[l] = 1.0 / va[l];
            __syncthreads();
            H_fresh = true;
            inner = 0;
          }
        }
        if (maxed || nanfail) break;
      }
      if (overflow) { jerr = -10000 - (lidx + 1); state = ST_DONE; break; }
      if (maxed) { jerr = -(lidx + 1); state = ST_DONE; break; }
      if (nanfail) { jerr = 9998; state = ST_DONE; break; }

      // ---- re-linearise: f = az + sum_l a_l x~_l, w = q exp(f), wr = t - w ----
      lap(3);
      crumb(8);
      __syncthreads();
      double c0p[1] = {0.0};
      for (int l = tid; l < nin; l += T) {
        const double b = aact[l] * xia[l];
        gq2[l] = b;  // scratch: gq2 is only live inside a block-mode sweep
        c0p[0] += b * xma[l];
      }
      bsum<1>(c0p, red, phase);
      const double c0 = az - c0p[0];
      const int nz_in = mlp ? compact_nonzero(gq2, nin) : nin;
      ncols += nz_in;
      double acc[4] = {0.0, 0.0, 0.0, 0.0};  // sum w, sum t f, sum wr, held-out deviance
      for (int i = 2 * tid; i < ldx; i += 2 * T) {
        double f0 = c0, f1 = c0;
        if (mlp) {
          const double2 fc2 = rows_comb(Xs + i, ldx, cks, gq2, nz_in, f0, f1);
          f0 = fc2.x;
          f1 = fc2.y;
        } else
#pragma unroll 4
        for (int l = 0; l < nin; ++l) {
          const double b = gq2[l];
          if (b != 0.0) {
            const double2 x = ldg2(Xs + (size_t)ia[l] * ldx + i);
            f0 += b * x.x;
            f1 += b * x.y;
          }
        }
        double wn[2] = {0.0, 0.0}, rn[2] = {0.0, 0.0};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ii = i + e;
          if (ii < n) {
            const int f = fo[ii];
            const double fe = e ? f1 : f0;
            if (f != 0 && f != fit) {
              const double yi = y[ii];
              const double t = qv * yi;
              const double fc = fabs(fe) < fmax_ ? fe : copysign(fmax_, fe);
              wn[e] = qv * exp(fc);
              rn[e] = t - wn[e];
              acc[0] += wn[e];
              acc[1] += t * fe;
              acc[2] += rn[e];
            } else if (f != 0 && is_fold) {  // held-out row of this fold
              const double yi = y[ii];
              const double devy = yi == 0.0 ? 0.0 : yi * log(yi) - yi;
              acc[3] += 2.0 * (devy - (yi * fe - exp(fe)));
            }
          }
        }
        *reinterpret_cast<double2*>(wv + i) = make_double2(wn[0], wn[1]);
        *reinterpret_cast<double2*>(wrv + i) = make_double2(rn[0], rn[1]);
      }
      bsum<4>(acc, red, phase);
      v0 = acc[0];
      sumr = acc[2];
      cur_dev = (acc[1] - v0 - dv0) / dvr;
      cur_cv = acc[3];
      // ---- outer convergence (fishnet1: intercept and every active coefficient) ----
      double chg = v0 * (az - az0) * (az - az0);
      for (int l = tid; l < nin; l += T) {
        const double d = aact[l] - as_[l];
        chg = fmax(chg, va[l] * d * d);
      }
      chg = bmax(chg, red, phase);
      lap(6);
      if (!(chg >= shr)) {
        state = ST_WAIT_KKT;
        break;
      }
      // the step just taken says how far the working weights moved: a large one makes the kept
      // model a poor chord (coarse lambda steps: the iteration then oscillates up to maxit)
      big_step = chg > A.prm.chord_tol * shr;
    }
  }

  // ---------------- write back ----------------
  __syncthreads();
  crumb(9);
  if (STAGED) {
    for (int i = 2 * tid; i < ldx; i += 2 * T) {
      __stcs(reinterpret_cast<double2*>(Wg + i), *reinterpret_cast<const double2*>(wv + i));
      __stcs(reinterpret_cast<double2*>(WRg + i), *reinterpret_cast<const double2*>(wrv + i));
    }
  }
  for (int l = tid; l < nin && l < kNX; l += T) {
    IAg[l] = ia[l];
    AAg[l] = aact[l];
  }
  if (state != ST_DONE) {
    if (tid == 0) {
      const int pos = atomicAdd(counter, 1);
      nextlist[pos] = pid;
      ish[40] = pos;
    }
    __syncthreads();
    const int pos = ish[40];
    double* r = A.ws.R + (size_t)pos * ldx;
    for (int i = 2 * tid; i < ldx; i += 2 * T)
      *reinterpret_cast<double2*>(r + i) = *reinterpret_cast<const double2*>(wrv + i);
    if (tid == 0) PR->slot = pos;
  }
  if (tid == 0 && hslot >= 0) {
    __threadfence();
    atomicExch(A.ws.hflags + hslot, 0);
  }
  if (tid == 0) {
    atomicAdd(A.ws.colcount + 4, gram_fma);
    for (int k = 0; k < 16; ++k) atomicAdd(A.ws.phasecyc + k, (unsigned long long)tph[k]);
    atomicAdd(A.ws.colcount, (unsigned long long)ncols);
    const unsigned long long cyc = (unsigned long long)(clock64() - t_start);
    atomicAdd(A.ws.colcount + 1, cyc);   // sum of CTA lifetimes
    atomicMax(A.ws.colcount + 2, cyc);   // longest CTA of the block of rounds (straggler)
    atomicAdd(A.ws.colcount + 3, (unsigned long long)(state == ST_WAIT_KKT || state == ST_DONE));
    PR->state = state; PR->lidx = lidx; PR->lmu = lmu; PR->nin = nin; PR->nS = nS;
    PR->jerr = jerr; PR->nlp = nlp; PR->ngrad += 1;
    PR->al = al; PR->al0 = al0; PR->az = az; PR->v0 = v0; PR->sumr = sumr;
    PR->cur_dev = cur_dev; PR->cur_cv = cur_cv;
  }
  crumb(10);
}

// ---------------------------------------------------------------------------------------------
// finalize: cv.glmnet statistics + lambda.min + prediction for all cells (one CTA per gene)
// ---------------------------------------------------------------------------------------------
struct FinalArgs {
  RoundArgs R;
  const double* lmin_user;  // path variant: s = lambda.min per gene
  double* MU;               // n x B
  double* out4;             // 4 x B: lambda.max, lambda.min, sd.cv, index
  int* status;              // B: 0 ok, 1 sd(y)==0, 2 fit failed (mean prediction)
  int* stats;               // 4 x B: lmu, jerr, sweeps, gradient rounds (summed over fits)
  double* cvout;            // optional 3 x kNLam x B: lambda, cvm, cvsd
};

__global__ void __launch_bounds__(256) lasso_finalize_kernel(FinalArgs F) {
  __shared__ double cvm[kNLam], cvsd[kNLam], bco[kNX];
  __shared__ int sh_i[4];
  __shared__ double sh_d[4];
  const RoundArgs& A = F.R;
  const int g = blockIdx.x, tid = threadIdx.x, NF = A.prm.NF, n = A.n, ldx = A.ldx;
  const LassoProb* P0 = A.ws.prob + (size_t)g * NF;
  const double* lam = A.ws.lam + (size_t)g * kNLam;
  double* mu = F.MU + (size_t)g * n;
  int st = 0;
  if (A.ws.gstat[g]) st = 1;
  int L = P0->lmu;
  if (!st) {
    // a fold fit may have been one lambda past the master's path when the master stopped: a fatal
    // error there (lidx >= L) concerns a solution cv.glmnet never asks for
    for (int f = 0; f < NF; ++f)
      if ((P0[f].jerr > 0 && (f == 0 || P0[f].lidx < L)) || P0[f].lmu < 1) st = 2;
  }
  if (F.stats && tid == 0) {
    int nlp = 0, ngr = 0;
    for (int f = 0; f < NF; ++f) { nlp += P0[f].nlp; ngr += P0[f].ngrad; }
    F.stats[4 * g + 0] = L; F.stats[4 * g + 1] = P0->jerr; F.stats[4 * g + 2] = nlp; F.stats[4 * g + 3] = ngr;
  }
  if (st) {
    const double m = A.ws.gmean[g];
    for (int i = tid; i < n; i += 256) mu[i] = m;
    if (tid == 0) {
      F.out4[4 * g + 0] = 0; F.out4[4 * g + 1] = 0; F.out4[4 * g + 2] = 0; F.out4[4 * g + 3] = -1;
      F.status[g] = st;
    }
    return;
  }
  int left, right;
  double frac = 1.0;
  if (A.prm.auto_lambda) {
    const int nfolds = A.prm.nfolds;
    for (int l = tid; l < L; l += 256) {
      double wsum = 0, s = 0;
      for (int f = 0; f < nfolds; ++f) {
        const int lf = l < P0[f + 1].lmu ? l : P0[f + 1].lmu - 1;  // short fold path: last column
        const double c = A.ws.cvdev[((size_t)g * kNLam + lf) * kMaxFolds + f];
        s += P0[f + 1].ncv * c;
        wsum += P0[f + 1].ncv;
      }
      const double m = s / wsum;
      double s2 = 0;
      for (int f = 0; f < nfolds; ++f) {
        const int lf = l < P0[f + 1].lmu ? l : P0[f + 1].lmu - 1;
        const double d = A.ws.cvdev[((size_t)g * kNLam + lf) * kMaxFolds + f] - m;
        s2 += P0[f + 1].ncv * d * d;
      }
      cvm[l] = m;
      cvsd[l] = sqrt(s2 / wsum / (nfolds - 1));
    }
    __syncthreads();
    if (tid == 0) {
      double best = INFINITY;
      for (int l = 0; l < L; ++l) if (cvm[l] < best) best = cvm[l];
      int imin = -1;
      for (int l = 0; l < L; ++l) if (cvm[l] <= best) { imin = l; break; }  // largest lambda at the minimum
      sh_i[0] = imin;
      if (imin >= 0) {
        F.out4[4 * g + 0] = lam[0];
        F.out4[4 * g + 1] = lam[imin];
        F.out4[4 * g + 2] = (cvm[0] - cvm[imin]) / cvsd[imin];
        F.out4[4 * g + 3] = imin;
      }
    }
    if (F.cvout) {
      double* co = F.cvout + (size_t)g * 3 * kNLam;
      for (int l = tid; l < kNLam; l += 256) {
        co[l] = l < L ? lam[l] : NAN;
        co[kNLam + l] = l < L ? cvm[l] : NAN;
        co[2 * kNLam + l] = l < L ? cvsd[l] : NAN;
      }
    }
    __syncthreads();
    left = right = sh_i[0];
    if (left < 0) {  // every cvm is NaN: cv.glmnet cannot pick -> treated as a failed fit
      const double m = A.ws.gmean[g];
      for (int i = tid; i < n; i += 256) mu[i] = m;
      if (tid == 0) {
        F.out4[4 * g + 0] = 0; F.out4[4 * g + 1] = 0; F.out4[4 * g + 2] = 0; F.out4[4 * g + 3] = -1;
        F.status[g] = 2;
      }
      return;
    }
  } else {
    // predict.glmnet(s = lambda.min): glmnet's lambda.interp on the fitted sequence
    if (tid == 0) {
      double s = F.lmin_user[g];
      int lf = L - 1, rt = L - 1;
      double fr = 1.0;
      if (s > lam[0]) s = lam[0];
      if (s < lam[L - 1]) s = lam[L - 1];
      if (L > 1 && lam[0] > lam[L - 1]) {
        const double den = lam[0] - lam[L - 1];
        const double sfrac = (lam[0] - s) / den;
        lf = 0;
        rt = L - 1;
        for (int l = 0; l < L; ++l) if ((lam[0] - lam[l]) / den <= sfrac) lf = l;
        for (int l = L - 1; l >= 0; --l) if ((lam[0] - lam[l]) / den >= sfrac) rt = l;
        fr = lam[lf] == lam[rt] ? 1.0 : (s - lam[rt]) / (lam[lf] - lam[rt]);
      }
      sh_i[0] = lf; sh_i[1] = rt; sh_d[0] = fr;
      F.out4[4 * g + 0] = lam[0];
      F.out4[4 * g + 1] = s;
      F.out4[4 * g + 2] = NAN;
      F.out4[4 * g + 3] = rt;
    }
    __syncthreads();
    left = sh_i[0]; right = sh_i[1]; frac = sh_d[0];
  }
  // coefficients on the Xs scale: b_l = a_l * xi_l ; intercept shift sum b_l xm_l
  const size_t oL = (size_t)g * kNLam + left, oR = (size_t)g * kNLam + right;
  const int ninL = A.ws.ninp[oL], ninR = A.ws.ninp[oR];
  const int nmax = ninL > ninR ? ninL : ninR;
  const size_t pid0 = (size_t)g * NF;
  const int* IAg = A.ws.IA + pid0 * kNX;
  const double* XMc = A.ws.XM + pid0 * A.p_pad;
  const double* XIc = A.ws.XI + pid0 * A.p_pad;
  __shared__ double red[512];
  int phase = 0;
  double sh[1] = {0.0};
  for (int l = tid; l < nmax; l += 256) {
    const double aL = l < ninL ? A.ws.cap[oL * kNX + l] : 0.0;
    const double aR = l < ninR ? A.ws.cap[oR * kNX + l] : 0.0;
    const int k = IAg[l];
    const double b = (frac * aL + (1.0 - frac) * aR) * XIc[k];
    bco[l] = b;
    sh[0] += b * XMc[k];
  }
  bsum<1>(sh, red, phase);
  const double a0 = frac * A.ws.a0p[oL] + (1.0 - frac) * A.ws.a0p[oR] - sh[0];
  __syncthreads();
  if (!(fabs(a0) < INFINITY)) {  // a non-finite solution is a failed fit: mean prediction
    const double m = A.ws.gmean[g];
    for (int i = tid; i < n; i += 256) mu[i] = m;
    if (tid == 0) {
      F.out4[4 * g + 0] = 0; F.out4[4 * g + 1] = 0; F.out4[4 * g + 2] = 0; F.out4[4 * g + 3] = -1;
      F.status[g] = 2;
    }
    return;
  }
  for (int i = tid; i < n; i += 256) {
    double f = a0;
    for (int l = 0; l < nmax; ++l) {
      const double b = bco[l];
      if (b != 0.0) f += b * __ldg(A.Xs + (size_t)IAg[l] * ldx + i);
    }
    mu[i] = exp(f);
  }
  if (tid == 0) F.status[g] = 0;
}

template <int T, bool STAGED, bool XBUF>
cudaError_t launch_round_k(const RoundArgs& A, const int* list, int* nextlist, int* counter,
                           int grid, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(lasso_round_kernel<T, STAGED, XBUF>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  lasso_round_kernel<T, STAGED, XBUF><<<grid, T, smem, s>>>(A, list, nextlist, counter);
  return cudaGetLastError();
}

template <int T>
cudaError_t launch_round_t(const RoundArgs& A, const int* list, int* nextlist, int* counter,
                           int grid, int mode, size_t smem, cudaStream_t s) {
  if (mode == 2) return launch_round_k<T, true, true>(A, list, nextlist, counter, grid, smem, s);
  if (mode == 1) return launch_round_k<T, true, false>(A, list, nextlist, counter, grid, smem, s);
  return launch_round_k<T, false, false>(A, list, nextlist, counter, grid, smem, s);
}

template <int T, bool STAGED, int MINB>
cudaError_t launch_gram_k(const RoundArgs& A, const int* list, int* nextlist, int* counter, int grid,
                          size_t smem, cudaStream_t s, int cluster = 1) {
  cudaError_t e = cudaFuncSetAttribute(lasso_round_gram_kernel<T, STAGED, MINB>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (cluster > 1 && grid % cluster == 0) {
    // one cluster per gene: the fits of a gene are co-scheduled (same GPC, same instant), nothing else
    // of the cluster machinery is used
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int* cl = list;
    return cudaLaunchKernelEx(&cfg, lasso_round_gram_kernel<T, STAGED, MINB>, A, cl, nextlist, counter);
  }
  lasso_round_gram_kernel<T, STAGED, MINB><<<grid, T, smem, s>>>(A, list, nextlist, counter);
  return cudaGetLastError();
}

// Gene-adjacent problem list.  `lst` holds the live problems in completion order of the last round;
// the fits of a gene read the same design columns (their strong / active sets nearly coincide), so
// they are launched next to each other: the second to sixth reader of a column tile finds it in L2
// instead of HBM.  Genes keep the order in which their LAST fit finished (the kernel takes the list
// from the back: long-running genes first).  aligned: every gene gets NF slots, -1 = fit done.
void order_by_gene(std::vector<int>& lst, int NF, bool aligned) {
  const int live = (int)lst.size();
  std::vector<std::pair<int, int>> key(live);  // (gene's last completion position, pid)
  {
    std::vector<std::pair<int, int>> bygene(live);
    for (int i = 0; i < live; ++i) bygene[i] = {lst[i] / NF, i};
    std::sort(bygene.begin(), bygene.end());
    for (int a = 0; a < live;) {
      int b = a;
      while (b < live && bygene[b].first == bygene[a].first) ++b;
      const int last = bygene[b - 1].second;  // positions ascend inside a gene
      for (int i = a; i < b; ++i) key[i] = {last, lst[bygene[i].second]};
      a = b;
    }
  }
  std::sort(key.begin(), key.end());
  std::vector<int> out;
  out.reserve(aligned ? (size_t)live * 2 : (size_t)live);
  for (int a = 0; a < live;) {
    int b = a;
    while (b < live && key[b].first == key[a].first) ++b;
    if (aligned) {
      const size_t base = out.size();
      out.resize(base + NF, -1);
      for (int i = a; i < b; ++i) out[base + key[i].second % NF] = key[i].second;
    } else {
      for (int i = a; i < b; ++i) out.push_back(key[i].second);
    }
    a = b;
  }
  lst.swap(out);
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

int lasso_max_block(int NF, const DesignState& D, const LassoTuning& tune) {
  // problems in flight per block: enough waves that a round does not end on a part-filled one
  // (12,288 = 41 waves of 296 resident CTAs), bounded by a third of the free device memory
  int tot = 12288;
  if (tune.problems > 0) tot = tune.problems;
  const double per_problem = (double)D.ldx * 8 * 3 + (double)D.p_pad * (8 * 3 + 4 * 3) + (double)kNX * 28;
  const double per_gene = NF * per_problem + (double)kNLam * (kNX * 8 + kMaxFolds * 8 + 32);
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
    const double cap = ((double)free_b / 3.0 - (double)kHSlots * (kNX * kNX + kGramItems * 256) * 8) / per_gene;
    if (cap * NF < tot) tot = (int)(cap > 1.0 ? cap * NF : NF);
  }
  int b = tot / NF;
  return b < 1 ? 1 : b;
}

// Runs one block of B genes.  All pointers are device pointers.
cudaError_t lasso_run(const DesignState& D, LassoWs& ws, const LassoParams& prm, const double* Yn,
                      const int* fold, const int* self, const double* lam_user,
                      const int* nlam_user, const double* lmin_user, double* MU, double* out4,
                      int* status, int* stats, double* cvout, cudaStream_t s, long long* launches,
                      long long* gemm_cols) {
  const int B = prm.B, NF = prm.NF, P = B * NF;
  if (B <= 0) return cudaSuccess;
  const int n = D.n, ldx = D.ldx, p_pad = D.p_pad;
  const int P_pad = (P + 63) / 64 * 64;
  // ---- carve the workspace ----
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t oR = take((size_t)ldx * P_pad * 8), oW = take((size_t)ldx * P_pad * 8),
               oWR = take((size_t)ldx * P_pad * 8), oG = take((size_t)p_pad * P_pad * 8),
               oXM = take((size_t)p_pad * P_pad * 8), oXI = take((size_t)p_pad * P_pad * 8),
               oMM = take((size_t)p_pad * P * 4), oSI = take((size_t)p_pad * P * 4),
               oIA = take((size_t)kNX * P * 4), oAA = take((size_t)kNX * P * 8), oAS = take((size_t)kNX * P * 8),
               oPR = take(sizeof(LassoProb) * (size_t)P), oLam = take((size_t)kNLam * B * 8),
               oNl = take((size_t)B * 4), oLf = take((size_t)B * 8), oGs = take((size_t)B * 4),
               oGm = take((size_t)B * 8), oA0 = take((size_t)kNLam * B * 8),
               oDv = take((size_t)kNLam * B * 8), oNi = take((size_t)kNLam * B * 4), oNp = take((size_t)kNLam * B * 4),
               oCap = take((size_t)kNX * kNLam * B * 8),
               oCv = take((size_t)kNLam * kMaxFolds * B * 8), oL0 = take((size_t)P * 4),
               oL1 = take((size_t)P * 4), oCt = take(64), oCc = take(64),
               oIn = take((size_t)p_pad * P * 4), oHf = take((size_t)kHSlots * 4),
               oHs = take((size_t)kHSlots * kNX * kNX * 8),
               oHp = take((size_t)kHSlots * kGramItems * 256 * 8), oPh = take(512);
  cudaError_t e;
  if (off > ws.cap_bytes) {
    if (ws.base) {
      if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
      cudaFree(ws.base);
      ws.base = nullptr;
      ws.cap_bytes = 0;
    }
    if ((e = cudaMalloc(&ws.base, off)) != cudaSuccess) return e;
    ws.cap_bytes = off;
  }
  if (!ws.h_counter && (e = cudaMallocHost(&ws.h_counter, 64)) != cudaSuccess) return e;
  unsigned char* b = ws.base;
  ws.R = (double*)(b + oR); ws.W = (double*)(b + oW); ws.WR = (double*)(b + oWR);
  ws.G = (double*)(b + oG); ws.XM = (double*)(b + oXM); ws.XI = (double*)(b + oXI);
  ws.MM = (int*)(b + oMM); ws.SIDX = (int*)(b + oSI); ws.IA = (int*)(b + oIA);
  ws.AACT = (double*)(b + oAA); ws.AS = (double*)(b + oAS); ws.prob = (LassoProb*)(b + oPR); ws.lam = (double*)(b + oLam);
  ws.nlam_g = (int*)(b + oNl); ws.msnap = (int*)(b + oLf); ws.gstat = (int*)(b + oGs);
  ws.gmean = (double*)(b + oGm); ws.a0p = (double*)(b + oA0); ws.devp = (double*)(b + oDv);
  ws.ninp = (int*)(b + oNi); ws.nlpp = (int*)(b + oNp); ws.cap = (double*)(b + oCap); ws.cvdev = (double*)(b + oCv);
  ws.list0 = (int*)(b + oL0); ws.list1 = (int*)(b + oL1); ws.counters = (int*)(b + oCt);
  ws.colcount = (unsigned long long*)(b + oCc);
  ws.INACT = (int*)(b + oIn); ws.hflags = (int*)(b + oHf); ws.Hs = (double*)(b + oHs); ws.HP = (double*)(b + oHp);
  ws.phasecyc = (unsigned long long*)(b + oPh);
  for (auto& ev : ws.ev)
    if (!ev && (e = cudaEventCreate(&ev)) != cudaSuccess) return e;

  RoundArgs A;
  A.Xs = D.Xs; A.Xf = ws.tune.f32 == 0 ? nullptr : D.Xf; A.xsd = D.xsd; A.n = n; A.ldx = ldx; A.p = D.p; A.p_pad = p_pad; A.P = P;
  A.prm = prm;
  const LassoTuning& tune = ws.tune;
  if (tune.dbg >= 0) A.prm.dbg = tune.dbg;
  if (tune.strong_slack >= 0) A.prm.strong_slack = tune.strong_slack;
  if (tune.chord_tol >= 0) A.prm.chord_tol = tune.chord_tol;
  if ((A.prm.dbg & 4096) && !ws.h_crumb) {
    if ((e = cudaHostAlloc(&ws.h_crumb, (size_t)kCrumbSlots * 32, cudaHostAllocMapped)) != cudaSuccess) return e;
    memset(ws.h_crumb, 0, (size_t)kCrumbSlots * 32);
    if ((e = cudaHostGetDevicePointer(&ws.d_crumb, ws.h_crumb, 0)) != cudaSuccess) return e;
  }
  A.round_no = 0;
  A.Yn = Yn; A.fold = fold; A.self = self; A.ws = ws;

  if ((e = cudaMemsetAsync(ws.counters, 0, 64, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ws.colcount, 0, 64, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ws.hflags, 0, (size_t)kHSlots * 4, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(ws.phasecyc, 0, 512, s)) != cudaSuccess) return e;
  cudaEventRecord(ws.ev[0], s);
  // the padding columns of the GEMM panel must hold finite numbers
  if (P_pad > P &&
      (e = cudaMemsetAsync(ws.R + (size_t)ldx * P, 0, (size_t)ldx * (P_pad - P) * 8, s)) != cudaSuccess)
    return e;
  lasso_setup_kernel<<<P, 256, 0, s>>>(A);
  if ((e = launch_gemm_tn_f64(D.Xs, ldx, ws.R, ldx, ws.XM, p_pad, D.p, P, ldx, s)) != cudaSuccess) return e;
  if ((e = launch_gemm_tn_f64(D.Xs2, ldx, ws.R, ldx, ws.XI, p_pad, D.p, P, ldx, s)) != cudaSuccess) return e;
  lasso_moments_kernel<<<dim3((p_pad + 255) / 256, P), 256, 0, s>>>(A);
  lasso_resid0_kernel<<<P, 256, 0, s>>>(A);
  if ((e = launch_gemm_tn_f64(D.Xs, ldx, ws.R, ldx, ws.G, p_pad, D.p, P, ldx, s)) != cudaSuccess) return e;
  *launches += 6;
  *gemm_cols += 3LL * P;
  if (prm.auto_lambda) {
    lasso_lambda_kernel<<<B, 256, 0, s>>>(A);
    *launches += 1;
  } else {
    if ((e = cudaMemcpyAsync(ws.lam, lam_user, (size_t)kNLam * B * 8, cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(ws.nlam_g, nlam_user, (size_t)B * 4, cudaMemcpyDeviceToDevice, s)) != cudaSuccess) return e;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;

  // ---- rounds ----
  // shared-memory plan: 2 = state + column ring (two CTAs per SM must still fit), 1 = state only,
  // 0 = everything through L2 (n too large to stage)
  const size_t smem_x = kFixedSmem + (size_t)4 * ldx * 8, smem_s = kFixedSmem + (size_t)2 * ldx * 8;
  const int mode = smem_x <= 113 * 1024 ? 2 : smem_s <= 227 * 1024 ? 1 : 0;
  const size_t smem = mode == 2 ? smem_x : mode == 1 ? smem_s : kFixedSmem;
  bool gram = prm.gram != 0;
  if (tune.gram >= 0) gram = tune.gram != 0;
  int T = ldx <= 2048 ? 128 : ldx <= 4096 ? 256 : ldx <= 8192 ? 512 : 1024;
  // Gram kernel: 256 threads and two CTAs per SM (128 registers each) while W / WR fit twice; above
  // that one CTA per SM of 512 threads.  Alternatives measured at n = 10,000 (profiles/README.md):
  // 256 threads x 255 registers (the streaming helpers then keep their load batches in flight, but
  // half the warps: -4 %), 1,024 threads x 64 registers (heavy spills: -20 %), 256 threads unstaged
  // with two CTAs per SM (-23 %).  SAVER_GRAM_THREADS / SAVER_GRAM_CTAS select them.
  int minb = ldx <= 4096 ? 2 : 1;
  if (gram) T = ldx <= 4096 ? 256 : 512;   // measured at n = 10,000: 512 x 128 regs 24.9 genes/s, 256 x 255 regs 23.9
  if (gram && tune.gram_threads > 0) T = tune.gram_threads;
  if (gram && T == 512) minb = 1;
  if (gram && tune.gram_ctas >= 1 && tune.gram_ctas <= 2 && T == 256) minb = tune.gram_ctas;
  if (!gram && tune.threads > 0) T = tune.threads;  // tuning experiments only
  if (T != 128 && T != 256 && T != 512 && T != 1024) return cudaErrorInvalidValue;
  if (gram && T != 256 && T != 512) return cudaErrorInvalidValue;
  // Gram kernel shared-memory plan: fixed state, then (staged) W / WR, then the packed Gram block
  // of the warp sweep in whatever is left of the CTA's share of the SM (two CTAs per SM up to 256
  // threads, one above).
  const size_t smem_g0 = kFixedSmem + kGramSmem, smem_gw = (size_t)2 * ldx * 8;
  const int ctas = minb;  // resident CTAs per SM the kernel variant is planned for
  size_t budget = (size_t)(228 * 1024 - ctas * 1024) / ctas / 16 * 16;
  if (budget > 227 * 1024) budget = 227 * 1024;
  bool gram_staged = smem_g0 + smem_gw <= budget;
  if (tune.gram_staged >= 0) gram_staged = gram_staged && tune.gram_staged != 0;
  if (!gram_staged && smem_g0 > budget) budget = 227 * 1024;
  size_t smem_gram = smem_g0 + (gram_staged ? smem_gw : 0);
  {
    size_t hs = smem_gram < budget ? (budget - smem_gram) / 16 * 16 : 0;
    if (tune.gram_hs >= 0 && hs > (size_t)tune.gram_hs) hs = (size_t)tune.gram_hs / 16 * 16;
    int mh = 0;
    while (mh < 256 && (size_t)(mh + 1) * (mh + 2) / 2 * 8 <= hs) ++mh;
    A.mh = mh;
    smem_gram += (size_t)mh * (mh + 1) / 2 * 8;
    smem_gram = (smem_gram + 15) / 16 * 16;
  }
  int cur = 0;
  cudaEventRecord(ws.ev[1], s);
  if ((e = cudaMemcpyAsync(ws.h_counter, ws.counters, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
  {
    float ms = 0;
    cudaEventElapsedTime(&ms, ws.ev[0], ws.ev[1]);
    ws.ms_setup += ms;
  }
  // The round kernels stream the design columns over and over while ~1 GB of per-problem panels
  // (gradients, moments, working weights) passes through L2 once per round: keep the design in the
  // persisting carve-out of L2 so that traffic does not evict it (hit rate of the column reads).
  bool l2_window = false;
  {
    int dev = 0, max_persist = 0, max_window = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    bool want = true;
    if (tune.l2_persist >= 0) want = tune.l2_persist != 0;
    if (want && max_persist > 0 && max_window > 0) {
      size_t bytes = (size_t)ldx * D.p * sizeof(double);
      if (bytes > (size_t)max_window) bytes = (size_t)max_window;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist) == cudaSuccess) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = const_cast<double*>(D.Xs);
        attr.accessPolicyWindow.num_bytes = bytes;
        const double ratio = (double)max_persist / (double)bytes;
        attr.accessPolicyWindow.hitRatio = ratio < 1.0 ? (float)ratio : 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        l2_window = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
      }
      cudaGetLastError();  // an unsupported attribute must not poison the launch checks below
    }
  }
  bool gemm_pending = false;
  int live = ws.h_counter[0];
  int rounds = 0;
  // problem order of a round: 0 = completion order, 1 = the fits of a gene adjacent, 2 = one
  // gene-aligned cluster of NF CTAs per gene (gram kernel, 2 <= NF <= 8)
  int order = gram && NF > 1 ? 1 : 0;
  if (tune.order >= 0) order = NF > 1 ? tune.order : 0;
  if (order == 2 && (!gram || NF > 8)) order = 1;
  std::vector<int> hlist;
  int grid = live;
  auto reorder = [&](int* dlist) -> cudaError_t {
    grid = live;
    if (order == 0 || live == 0) return cudaSuccess;
    hlist.resize(live);
    cudaError_t e2 = cudaMemcpyAsync(hlist.data(), dlist, (size_t)live * 4, cudaMemcpyDeviceToHost, s);
    if (e2 != cudaSuccess) return e2;
    if ((e2 = cudaStreamSynchronize(s)) != cudaSuccess) return e2;
    order_by_gene(hlist, NF, order == 2);
    grid = (int)hlist.size();
    return cudaMemcpyAsync(dlist, hlist.data(), (size_t)grid * 4, cudaMemcpyHostToDevice, s);
  };
  if ((e = reorder(ws.list0)) != cudaSuccess) return e;
  while (live > 0) {
    if (++rounds > 100000) return cudaErrorUnknown;
    int* list = cur ? ws.list1 : ws.list0;
    int* next = cur ? ws.list0 : ws.list1;
    int* cnt = ws.counters + (cur ? 0 : 1);
    if ((e = cudaMemsetAsync(cnt, 0, 4, s)) != cudaSuccess) return e;
    if (prm.auto_lambda) {
      lasso_snapshot_kernel<<<(B + 255) / 256, 256, 0, s>>>(A);
      *launches += 1;
    }
    cudaEventRecord(ws.ev[0], s);
    A.round_no = rounds;
    if (gram) {
      const size_t sm = smem_gram;
      const int cl = order == 2 ? NF : 1;
      if (T == 512) e = gram_staged ? launch_gram_k<512, true, 1>(A, list, next, cnt, grid, sm, s, cl)
                                    : launch_gram_k<512, false, 1>(A, list, next, cnt, grid, sm, s, cl);
      else if (minb == 2) e = gram_staged ? launch_gram_k<256, true, 2>(A, list, next, cnt, grid, sm, s, cl)
                                          : launch_gram_k<256, false, 2>(A, list, next, cnt, grid, sm, s, cl);
      else e = gram_staged ? launch_gram_k<256, true, 1>(A, list, next, cnt, grid, sm, s, cl)
                           : launch_gram_k<256, false, 1>(A, list, next, cnt, grid, sm, s, cl);
    } else
    switch (T) {
      case 128: e = launch_round_t<128>(A, list, next, cnt, live, mode, smem, s); break;
      case 256: e = launch_round_t<256>(A, list, next, cnt, live, mode, smem, s); break;
      case 512: e = launch_round_t<512>(A, list, next, cnt, live, mode, smem, s); break;
      default: e = launch_round_t<1024>(A, list, next, cnt, live, mode, smem, s); break;
    }
    if (e != cudaSuccess) return e;
    *launches += 1;
    cudaEventRecord(ws.ev[1], s);
    if ((e = cudaMemcpyAsync(ws.h_counter, cnt, 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) {
      if (ws.h_crumb) {  // the mapped breadcrumbs survive the fault: CTAs that never reached tag 10
        fprintf(stderr, "lasso_run: round %d (live %d, T %d) failed: %s\n", rounds, live, T, cudaGetErrorString(e));
        for (int c = 0; c < kCrumbSlots; ++c) {
          const int* r = ws.h_crumb + (size_t)c * 8;
          if (r[2] != 0 && r[2] != 10)
            fprintf(stderr, "  crumb slot %d: round %d pid %d tag %d lidx %d nin %d nS %d m %d ninact %d\n",
                    c, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
        }
      }
      return e;
    }
    {
      float ms = 0;
      cudaEventElapsedTime(&ms, ws.ev[0], ws.ev[1]);
      ws.ms_round += ms;
      if (gemm_pending) {
        cudaEventElapsedTime(&ms, ws.ev[2], ws.ev[3]);
        ws.ms_gemm += ms;
      }
      ws.rounds += 1;
    }
    live = ws.h_counter[0];
    if ((e = reorder(next)) != cudaSuccess) return e;
    cur ^= 1;
    gemm_pending = false;
    if (live > 0) {
      cudaEventRecord(ws.ev[2], s);
      if ((e = launch_gemm_tn_f64(D.Xs, ldx, ws.R, ldx, ws.G, p_pad, D.p, live, ldx, s)) != cudaSuccess) return e;
      cudaEventRecord(ws.ev[3], s);
      gemm_pending = true;
      *launches += 1;
      *gemm_cols += live;
    }
  }
  {
    unsigned long long cc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if ((e = cudaMemcpyAsync(cc, ws.colcount, 64, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    ws.cols_streamed += cc[0];
    ws.gram_fma += cc[4];
    unsigned long long ph[48];
    if ((e = cudaMemcpy(ph, ws.phasecyc, 48 * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    for (int k = 0; k < 48; ++k) ws.phase_cycles[k] += ph[k];
    ws.cta_cycles += cc[1];
    if (cc[2] > ws.max_cta_cycles) ws.max_cta_cycles = cc[2];
  }
  if (l2_window) {
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
  }
  FinalArgs F;
  F.R = A; F.lmin_user = lmin_user; F.MU = MU; F.out4 = out4; F.status = status; F.stats = stats;
  F.cvout = cvout;
  lasso_finalize_kernel<<<B, 256, 0, s>>>(F);
  *launches += 1;
  return cudaGetLastError();
}

// diagnostics: master path of gene g of the last block (a0, dev, nin, sweeps per lambda)
cudaError_t lasso_debug_path(const LassoWs& ws, int g, double* a0, double* dev, int* nin, int* nlp,
                             double* lam) {
  if (!ws.base) return cudaErrorInvalidValue;
  cudaError_t e;
  if ((e = cudaMemcpy(a0, ws.a0p + (size_t)g * kNLam, kNLam * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(dev, ws.devp + (size_t)g * kNLam, kNLam * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(nin, ws.ninp + (size_t)g * kNLam, kNLam * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(nlp, ws.nlpp + (size_t)g * kNLam, kNLam * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
  return cudaMemcpy(lam, ws.lam + (size_t)g * kNLam, kNLam * 8, cudaMemcpyDeviceToHost);
}

}  // namespace saver
