// estimate.cu -- saver_cuda_calc_estimate: the body of the reference's data-parallel loop for one
// block of genes, entirely on the device.
//
// Replaces R/calc_estimate.R:75-141 (the `foreach ... %dopar%` body of calc.estimate):
//   y = sweep(ix, 2, sf, "/")                                   :75   normalize_kernel
//   maxcor = calc.maxcor(x.est, t(y))  | rep(2, .)               :76-80 design.cu + gemm64.cu
//   pred.gene = (maxcor > cutoff) & (name %in% pred.gene.names)  :96
//   per gene: mean | expr.predict CV | est.lambda + expr.predict :100-129 lasso.cu
//   calc.post(ix[i,], mu, sf, scale.sf)                          :134  variance.cu
// The per-gene loop of the reference becomes three batched phases over the block.
#include <math.h>
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"
#include "ctx.h"

using namespace saver;

namespace {

__global__ void normalize_kernel(const double* __restrict__ X, const double* __restrict__ sf, int n,
                                 int B, double* __restrict__ Y) {
  const size_t tot = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    Y[i] = X[i] / sf[i % n];
}

// one CTA per gene: mean over the pred cells, sd (n-1) over all cells (two-pass, as stats::var)
__global__ void __launch_bounds__(256)
gene_stats_kernel(const double* __restrict__ Y, const int* __restrict__ mask, int n,
                  double* __restrict__ mean_pred, double* __restrict__ sd_all) {
  __shared__ double scratch[24];
  const int g = blockIdx.x;
  const double* y = Y + (size_t)g * n;
  double v[3] = {0.0, 0.0, 0.0};  // sum all, sum pred, count pred
  for (int i = threadIdx.x; i < n; i += 256) {
    const double yi = y[i];
    v[0] += yi;
    if (!mask || mask[i]) { v[1] += yi; v[2] += 1.0; }
  }
  block_sum<3>(v, scratch);
  double m = v[0] / n, t = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) t += y[i] - m;
  m += block_sum1(t, scratch) / n;
  double ss = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { const double d = y[i] - m; ss += d * d; }
  ss = block_sum1(ss, scratch);
  if (threadIdx.x == 0) {
    mean_pred[g] = v[2] > 0 ? v[1] / v[2] : 0.0;
    sd_all[g] = sqrt(ss / (n - 1));
  }
}

template <typename T>
__global__ void gather_cols_kernel(const T* __restrict__ src, const int* __restrict__ idx, int n,
                                   int Bp, T* __restrict__ dst) {
  const size_t tot = (size_t)n * Bp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[(size_t)idx[i / n] * n + i % n];
}

__global__ void scatter_cols_kernel(const double* __restrict__ src, const int* __restrict__ idx,
                                    int n, int Bp, double* __restrict__ dst) {
  const size_t tot = (size_t)n * Bp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    dst[(size_t)idx[i / n] * n + i % n] = src[i];
}

__global__ void fill_cols_kernel(double* dst, const double* scalars, int n, int B) {
  const size_t tot = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = scalars[i / n];
}

__global__ void mask_to_fold_kernel(const int* mask, int n, int B, int* fold) {
  const size_t tot = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    fold[i] = mask ? (mask[i % n] != 0) : 1;
}

// foldid (0 outside pred cells) from the host's per-gene fold ids and the pred-cell mask
__global__ void apply_mask_kernel(int* fold, const int* mask, int n, int B) {
  const size_t tot = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    if (!mask[i % n]) fold[i] = 0;
}

double lambda_ratio(double flmin, int nlam) {
  const double eqs = flmin > 1e-6 ? flmin : 1e-6;
  return nlam > 1 ? pow(eqs, (double)(1.0f / (float)(nlam - 1))) : 1.0;
}

// lambda.seq of R/expr_predict.R:62-63, sorted decreasing as glmnet does with user lambdas
int build_lambda_seq(double lmax, double lmin, double* lg) {
  int nl = 0;
  if (lmax > 0 && lmin > 0 && lmax == lmax && lmin == lmin) {
    const double from = log(lmax), to = log(lmin);
    int nstep = (int)floor((to - from) / -0.2 + 1e-10);
    if (nstep < 0) nstep = 0;
    if (nstep + 2 > kNLam) return -1;
    for (int i = 0; i <= nstep; ++i) lg[nl++] = exp(from + i * -0.2);
    lg[nl++] = lmin;
    for (int i = 1; i < nl; ++i) {
      const double v = lg[i];
      int j = i - 1;
      while (j >= 0 && lg[j] < v) { lg[j + 1] = lg[j]; --j; }
      lg[j + 1] = v;
    }
  } else {
    lg[nl++] = lmax;
  }
  return nl;
}

const int kGrid = 148 * 4;

}  // namespace

namespace saver {
int build_lambda_seq_host(double lmax, double lmin, double* lg) { return build_lambda_seq(lmax, lmin, lg); }
double lambda_ratio_host(double flmin, int nlam) { return lambda_ratio(flmin, nlam); }
}  // namespace saver

static int calc_estimate_one(saver_cuda_handle h, const double* X, int n_cells, int B, const double* sf,
                                        double scale_sf, const int* self_col, const int* is_pred,
                                        const int* pred_mask, const int* foldid, int nfolds,
                                        int mode, int calc_maxcor, double cutoff,
                                        const double* coefs, int dfmax, int nlambda, double thresh,
                                        double* est, double* se, double* info6, double* mu_out,
                                        int* status_out, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!X || !sf || !est || !info6 || B < 0 || mode < 0 || mode > 2)
    return fail(h, SAVER_ERR_ARG, "calc_estimate: bad arguments");
  if ((mode != 0 || calc_maxcor) && !h->design.Xs)
    return fail(h, SAVER_ERR_STATE, "calc_estimate: design not set");
  if (h->design.Xs && (mode != 0 || calc_maxcor) && h->design.n != n_cells)
    return fail(h, SAVER_ERR_ARG, "calc_estimate: block has %d cells, design has %d", n_cells, h->design.n);
  if (mode == 1 && (!foldid || nfolds < 3 || nfolds > kMaxFolds))
    return fail(h, SAVER_ERR_ARG, "calc_estimate: CV needs foldid and 3 <= nfolds <= 16");
  if (mode == 2 && !coefs) return fail(h, SAVER_ERR_ARG, "calc_estimate: fixed-lambda needs coefs");
  // same limits as saver_cuda_predict_cv: the lambda tables hold kNLam entries, the ever-active list
  // kNX (glmnet's pmax = 2 * dfmax + 20 must fit, otherwise the path would truncate differently)
  if (mode != 0 && (nlambda < 1 || nlambda > kNLam || dfmax < 1 || 2 * (long long)dfmax + 20 > kNX ||
                    !(thresh > 0)))
    return fail(h, SAVER_ERR_ARG,
                "calc_estimate: need 1 <= nlambda <= %d, 1 <= dfmax <= %d, thresh > 0 (got %d, %d, %g)",
                kNLam, (kNX - 20) / 2, nlambda, dfmax, thresh);
  if (B == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  const bool dev = device_ptrs != 0;
  cudaStream_t s = h->stream;
  const DesignState& D = h->design;
  const size_t n = (size_t)n_cells;
  if (n_cells < 2) return fail(h, SAVER_ERR_ARG, "calc_estimate: n < 2");
  const size_t nb = n * B;

  cudaEvent_t ev[3];
  for (auto& e : ev) CU(cudaEventCreate(&e));
  struct EvGuard { cudaEvent_t* e; ~EvGuard() { for (int i = 0; i < 3; ++i) cudaEventDestroy(e[i]); } } evg{ev};

  InArr<double> dX, dsf;
  InArr<int> dSelf, dMask, dFold;
  CU(dX.bind(X, nb, dev, s));
  CU(dsf.bind(sf, n, dev, s));
  CU(dMask.bind(pred_mask, n, dev, s));
  OutArr<double> oE, oS, oMu;
  CU(oE.bind(est, nb, dev, s));
  CU(oS.bind(se, nb, dev, s));
  DevBuf<double> Yn, MUbuf, d_mc, d_mean, d_sd, info_tmp;
  CU(Yn.alloc(nb, s));
  double* MU;
  if (mu_out) { CU(oMu.bind(mu_out, nb, dev, s)); MU = oMu.d; }
  else { CU(MUbuf.alloc(nb, s)); MU = MUbuf.p; }
  CU(d_mc.alloc(B, s));
  CU(d_mean.alloc(B, s));
  CU(d_sd.alloc(B, s));
  CU(info_tmp.alloc((size_t)kPostInfo * B, s));

  CU(cudaEventRecord(ev[0], s));
  normalize_kernel<<<kGrid, 256, 0, s>>>(dX.d, dsf.d, (int)n, B, Yn.p);
  gene_stats_kernel<<<B, 256, 0, s>>>(Yn.p, dMask.d, (int)n, d_mean.p, d_sd.p);
  fill_cols_kernel<<<kGrid, 256, 0, s>>>(MU, d_mean.p, (int)n, B);
  CU(cudaGetLastError());
  h->launches += 3;

  std::vector<double> h_mc(B, 2.0), h_sd(B), h_lmax(B, 0.0), h_lmin(B, 0.0), h_sdcv(B, 0.0);
  std::vector<int> h_self(B, -1), h_pred(B, 1), h_status(B, 0);
  if (mode != 0 || calc_maxcor) {
    if (self_col) {
      if (dev) CU(cudaMemcpyAsync(h_self.data(), self_col, B * sizeof(int), cudaMemcpyDeviceToHost, s));
      else h_self.assign(self_col, self_col + B);
    }
    if (is_pred) {
      if (dev) CU(cudaMemcpyAsync(h_pred.data(), is_pred, B * sizeof(int), cudaMemcpyDeviceToHost, s));
      else h_pred.assign(is_pred, is_pred + B);
    }
    if (calc_maxcor) {
      // calc.maxcor over the whole block (R/calc_estimate.R:77)
      DevBuf<int> dSelfAll;
      CU(dSelfAll.alloc(B, s));
      if (dev && self_col) CU(cudaMemcpyAsync(dSelfAll.p, self_col, B * sizeof(int), cudaMemcpyDeviceToDevice, s));
      else {
        if (dev) CU(cudaStreamSynchronize(s));  // h_self must have landed
        CU(cudaMemcpyAsync(dSelfAll.p, h_self.data(), B * sizeof(int), cudaMemcpyHostToDevice, s));
      }
      const int chunk = 1024;
      DevBuf<double> R, Cm;
      CU(R.alloc((size_t)D.ldx * chunk, s));
      CU(Cm.alloc((size_t)D.p_pad * chunk, s));
      for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nbk = B - b0 < chunk ? B - b0 : chunk;
        CU(launch_center_scale(Yn.p + n * b0, D.n, D.ldx, nbk, R.p, s));
        CU(launch_gemm_tn_f64(D.Xs, D.ldx, R.p, D.ldx, Cm.p, D.p_pad, D.p, nbk, D.ldx, s));
        CU(launch_colmax(Cm.p, D.p_pad, D.p, nbk, dSelfAll.p + b0, d_mc.p + b0, s));
        h->launches += 3;
      }
      CU(cudaMemcpyAsync(h_mc.data(), d_mc.p, B * sizeof(double), cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));  // dSelfAll's upload source and h_mc
    }
    CU(cudaMemcpyAsync(h_sd.data(), d_sd.p, B * sizeof(double), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));

    // pred.gene <- (maxcor > cutoff) & (x.names %in% pred.gene.names)   (R/calc_estimate.R:96)
    std::vector<int> idx;
    for (int g = 0; g < B && mode != 0; ++g)
      if (h_mc[g] > cutoff && h_pred[g]) idx.push_back(g);
    const int Bp = (int)idx.size();
    if (Bp > 0) {
      std::vector<int> selfp(Bp);
      for (int j = 0; j < Bp; ++j) selfp[j] = h_self[idx[j]];
      DevBuf<int> dIdx, dSelfP, dFoldP, dStat;
      DevBuf<double> Yp, MUp, fit4;
      CU(dIdx.alloc(Bp, s)); CU(dSelfP.alloc(Bp, s)); CU(dStat.alloc(Bp, s));
      CU(Yp.alloc(n * Bp, s)); CU(MUp.alloc(n * Bp, s)); CU(fit4.alloc((size_t)4 * Bp, s));
      CU(dFoldP.alloc(n * Bp, s));
      CU(cudaMemcpyAsync(dIdx.p, idx.data(), Bp * sizeof(int), cudaMemcpyHostToDevice, s));
      CU(cudaMemcpyAsync(dSelfP.p, selfp.data(), Bp * sizeof(int), cudaMemcpyHostToDevice, s));
      gather_cols_kernel<double><<<kGrid, 256, 0, s>>>(Yn.p, dIdx.p, (int)n, Bp, Yp.p);
      h->launches += 1;
      LassoParams prm;
      prm.dfmax = dfmax; prm.thr = thresh; prm.maxit = 100000;
  prm.gram = h->opt_gram; prm.strong_slack = h->opt_strong_slack; prm.dbg = h->opt_dbg; prm.inner_scale = h->opt_inner_scale;
      std::vector<double> lam;
      std::vector<int> nlam;
      std::vector<double> lminp;
      DevBuf<double> dLam, dLmin;
      DevBuf<int> dNlam;
      if (mode == 1) {
        CU(dFold.bind(foldid, nb, dev, s));
        gather_cols_kernel<int><<<kGrid, 256, 0, s>>>(dFold.d, dIdx.p, (int)n, Bp, dFoldP.p);
        if (dMask.d) apply_mask_kernel<<<kGrid, 256, 0, s>>>(dFoldP.p, dMask.d, (int)n, Bp);
        h->launches += 2;
        prm.NF = nfolds + 1; prm.nfolds = nfolds; prm.auto_lambda = 1; prm.nlam = nlambda;
        prm.alf_wide = lambda_ratio(0.01, nlambda);
        prm.alf_tall = lambda_ratio(1e-4, nlambda);
      } else {
        // est.lambda (R/utils.R:129-134): lambda.max = r sd(y) sqrt((n-1)/n), lambda.min = ...
        lam.assign((size_t)kNLam * Bp, 0.0);
        nlam.assign(Bp, 0);
        lminp.assign(Bp, 0.0);
        for (int j = 0; j < Bp; ++j) {
          const int g = idx[j];
          const double r = h_mc[g];
          const double lmax = r * h_sd[g] * sqrt(((double)n - 1.0) / (double)n);
          double arg = coefs[0] + coefs[1] * r;
          if (!(arg > 0)) arg = 0;
          const double lmin = lmax * exp(-sqrt(arg));
          h_lmax[g] = lmax;
          h_lmin[g] = lmin;
          lminp[j] = lmin;
          const int nl = build_lambda_seq(lmax, lmin, lam.data() + (size_t)kNLam * j);
          if (nl < 0) return fail(h, SAVER_ERR_ARG, "calc_estimate: lambda sequence too long");
          nlam[j] = nl;
        }
        CU(dLam.alloc(lam.size(), s)); CU(dNlam.alloc(Bp, s)); CU(dLmin.alloc(Bp, s));
        CU(cudaMemcpyAsync(dLam.p, lam.data(), lam.size() * 8, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dNlam.p, nlam.data(), Bp * 4, cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(dLmin.p, lminp.data(), Bp * 8, cudaMemcpyHostToDevice, s));
        mask_to_fold_kernel<<<kGrid, 256, 0, s>>>(dMask.d, (int)n, Bp, dFoldP.p);
        h->launches += 1;
        prm.NF = 1; prm.nfolds = 0; prm.auto_lambda = 0; prm.nlam = 0;
        prm.alf_wide = prm.alf_tall = 1.0;
      }
      CU(cudaGetLastError());
      const int chunk = lasso_max_block(prm.NF, D, h->lasso.tune);
      for (int b0 = 0; b0 < Bp; b0 += chunk) {
        prm.B = Bp - b0 < chunk ? Bp - b0 : chunk;
        CU(lasso_run(D, h->lasso, prm, Yp.p + n * b0, dFoldP.p + n * b0, dSelfP.p + b0,
                     mode == 2 ? dLam.p + (size_t)kNLam * b0 : nullptr,
                     mode == 2 ? dNlam.p + b0 : nullptr, mode == 2 ? dLmin.p + b0 : nullptr,
                     MUp.p + n * b0, fit4.p + (size_t)4 * b0, dStat.p + b0, nullptr, nullptr, s,
                     &h->launches, &h->gemm_cols));
      }
      scatter_cols_kernel<<<kGrid, 256, 0, s>>>(MUp.p, dIdx.p, (int)n, Bp, MU);
      CU(cudaGetLastError());
      h->launches += 1;
      std::vector<double> h_fit((size_t)4 * Bp);
      std::vector<int> h_st(Bp);
      CU(cudaMemcpyAsync(h_fit.data(), fit4.p, h_fit.size() * 8, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpyAsync(h_st.data(), dStat.p, Bp * 4, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      for (int j = 0; j < Bp; ++j) {
        const int g = idx[j];
        h_status[g] = h_st[j];
        if (mode == 1) {
          h_lmax[g] = h_fit[4 * j + 0];
          h_lmin[g] = h_fit[4 * j + 1];
          h_sdcv[g] = h_fit[4 * j + 2];
        } else {
          h_sdcv[g] = h_st[j] == 0 ? NAN : 0.0;  // NA on success, 0 from the fallbacks
        }
      }
    }
  }
  CU(cudaEventRecord(ev[1], s));

  // ---- calc.post for the whole block ----
  int stage;
  calc_post_smem_bytes((int)n, &stage);
  DevBuf<int> nz;
  if (stage == 0) CU(nz.alloc(nb, s));
  CalcPostArgs PA{dX.d, MU, dsf.d, (int)n, B, scale_sf, oE.d, oS.d, nullptr, nullptr, info_tmp.p,
                  nullptr, nz.p};
  CU(launch_calc_post(PA, s));
  h->launches += 1;
  CU(cudaEventRecord(ev[2], s));
  CU(oE.fetch(s));
  CU(oS.fetch(s));
  CU(oMu.fetch(s));
  CU(cudaStreamSynchronize(s));
  float ms_pred = 0, ms_var = 0;
  CU(cudaEventElapsedTime(&ms_pred, ev[0], ev[1]));
  CU(cudaEventElapsedTime(&ms_var, ev[1], ev[2]));
  h->ms_post += ms_var;
  // info6: six contiguous vectors of B (maxcor, lambda.max, lambda.min, sd.cv, pred.time, var.time)
  std::vector<double> info((size_t)6 * B);
  for (int g = 0; g < B; ++g) {
    info[0 * (size_t)B + g] = h_mc[g];
    info[1 * (size_t)B + g] = h_lmax[g];
    info[2 * (size_t)B + g] = h_lmin[g];
    info[3 * (size_t)B + g] = h_sdcv[g];
    info[4 * (size_t)B + g] = ms_pred * 1e-3 / B;
    info[5 * (size_t)B + g] = ms_var * 1e-3 / B;
  }
  if (dev) {
    CU(cudaMemcpyAsync(info6, info.data(), info.size() * 8, cudaMemcpyHostToDevice, s));
    if (status_out) CU(cudaMemcpyAsync(status_out, h_status.data(), B * 4, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
  } else {
    memcpy(info6, info.data(), info.size() * 8);
    if (status_out) memcpy(status_out, h_status.data(), B * 4);
  }
  return SAVER_OK;
}

// The public entry point.  A one-device handle runs the block as it is.  A multi handle
// (saver_cuda_create_multi) splits the block's genes into contiguous chunks, one per device, and runs
// them on one host thread each -- what calc.estimate does with its row chunks and PSOCK workers
// (R/calc_estimate.R:69-74,147-160); panels are one gene per column, so a chunk is a pointer offset.
extern "C" int saver_cuda_calc_estimate(saver_cuda_handle h, const double* X, int n_cells, int B, const double* sf,
                                        double scale_sf, const int* self_col, const int* is_pred,
                                        const int* pred_mask, const int* foldid, int nfolds,
                                        int mode, int calc_maxcor, double cutoff,
                                        const double* coefs, int dfmax, int nlambda, double thresh,
                                        double* est, double* se, double* info6, double* mu_out,
                                        int* status_out, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  const int ndev = 1 + (int)h->peers.size();
  if (ndev == 1 || B < 2 * ndev)
    return calc_estimate_one(h, X, n_cells, B, sf, scale_sf, self_col, is_pred, pred_mask, foldid, nfolds, mode,
                             calc_maxcor, cutoff, coefs, dfmax, nlambda, thresh, est, se, info6, mu_out,
                             status_out, device_ptrs);
  if (device_ptrs) return fail(h, SAVER_ERR_ARG, "calc_estimate: a multi-device handle takes host buffers");
  if (!X || !est || !info6 || n_cells < 2) return fail(h, SAVER_ERR_ARG, "calc_estimate: bad arguments");
  std::vector<saver_cuda_ctx*> ctx(1, h);
  ctx.insert(ctx.end(), h->peers.begin(), h->peers.end());
  std::vector<int> b0(ndev + 1), rc(ndev, SAVER_OK);
  for (int d = 0; d <= ndev; ++d) b0[d] = (int)((long long)B * d / ndev);
  std::vector<std::vector<double>> info(ndev);
  std::vector<std::thread> th;
  const size_t n = (size_t)n_cells;
  for (int d = 0; d < ndev; ++d) {
    const int o = b0[d], nb = b0[d + 1] - b0[d];
    info[d].assign((size_t)6 * nb, 0.0);
    th.emplace_back([=, &rc, &info]() {
      rc[d] = calc_estimate_one(ctx[d], X + n * o, n_cells, nb, sf, scale_sf, self_col ? self_col + o : nullptr,
                                is_pred ? is_pred + o : nullptr, pred_mask, foldid ? foldid + n * o : nullptr,
                                nfolds, mode, calc_maxcor, cutoff, coefs, dfmax, nlambda, thresh, est + n * o,
                                se ? se + n * o : nullptr, info[d].data(), mu_out ? mu_out + n * o : nullptr,
                                status_out ? status_out + o : nullptr, 0);
    });
  }
  for (auto& t : th) t.join();
  for (int d = 0; d < ndev; ++d) {
    if (rc[d] != SAVER_OK) return fail(h, rc[d], "device %d: %s", ctx[d]->device, ctx[d]->err);
    const int o = b0[d], nb = b0[d + 1] - b0[d];
    for (int k = 0; k < 6; ++k)
      for (int g = 0; g < nb; ++g) info6[(size_t)k * B + o + g] = info[d][(size_t)k * nb + g];
  }
  return SAVER_OK;
}
