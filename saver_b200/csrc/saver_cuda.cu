// saver_cuda.cu -- the C ABI of libsaver_cuda (include/saver_cuda.h).
// Marshals caller buffers to the device, launches the kernels, copies results back.
#include "../../include/saver_cuda.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "ctx.h"

using namespace saver;

namespace {

__global__ void fill_panel_kernel(double* dst, const double* scalars, int n, int B) {
  const size_t tot = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    dst[i] = scalars[i / n];
}

}  // namespace

extern "C" {

int saver_cuda_abi_version(void) { return SAVER_ABI_VERSION; }

int saver_cuda_create(int device, saver_cuda_handle* out) {
  if (!out) return SAVER_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return SAVER_ERR_CUDA;
  saver_cuda_ctx* h = new (std::nothrow) saver_cuda_ctx();
  if (!h) return SAVER_ERR_NOMEM;
  h->device = device;
  h->lasso.tune.from_env();  // developer overrides are read once, here
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return SAVER_ERR_CUDA;
  }
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  *out = h;
  return SAVER_OK;
}

// Several devices behind one handle: the in-process analogue of the reference's ncores = PSOCK
// workers (R/saver.R:120-138) for hosts that are ONE process (an R session).  The first device is
// the handle itself; the others are ordinary handles kept in `peers`.
int saver_cuda_create_multi(int ndev, const int* devs, saver_cuda_handle* out) {
  if (!out) return SAVER_ERR_ARG;
  *out = nullptr;
  if (ndev < 1 || !devs) return SAVER_ERR_ARG;
  saver_cuda_handle h = nullptr;
  int rc = saver_cuda_create(devs[0], &h);
  if (rc != SAVER_OK) return rc;
  for (int i = 1; i < ndev; ++i) {
    saver_cuda_handle q = nullptr;
    rc = saver_cuda_create(devs[i], &q);
    if (rc != SAVER_OK) {
      saver_cuda_destroy(h);
      return rc;
    }
    h->peers.push_back(q);
  }
  *out = h;
  return SAVER_OK;
}

int saver_cuda_device_count(saver_cuda_handle h) { return h ? 1 + (int)h->peers.size() : 0; }

int saver_cuda_destroy(saver_cuda_handle h) {
  if (!h) return SAVER_OK;
  for (saver_cuda_ctx* q : h->peers) saver_cuda_destroy(q);
  h->peers.clear();
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  h->design.release();
  h->lasso.release();
  cudaFree(h->csc_indptr); cudaFree(h->csc_indices); cudaFree(h->csc_data); cudaFree(h->csc_map);
  cudaFree(h->dn_x);
  cudaStreamDestroy(h->stream);
  delete h;
  return SAVER_OK;
}

const char* saver_cuda_last_error(saver_cuda_handle h) { return h ? h->err : "null handle"; }

int saver_cuda_synchronize(saver_cuda_handle h) {
  if (!h) return SAVER_ERR_ARG;
  for (saver_cuda_ctx* q : h->peers) {
    const int rc = saver_cuda_synchronize(q);
    if (rc != SAVER_OK) return fail(h, rc, "%s", q->err);
  }
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  return SAVER_OK;
}

int saver_cuda_stream(saver_cuda_handle h, void** stream) {
  if (!h || !stream) return SAVER_ERR_ARG;
  *stream = (void*)h->stream;
  return SAVER_OK;
}

long long saver_cuda_launch_count(saver_cuda_handle h) { return h ? h->launches : 0; }

static int one_gene(saver_cuda_handle h, int mode, int model, double param, const double* y,
                    const double* mu, int mu_len, const double* sf, int sf_len, int n,
                    double* out4, double* grid200) {
  if (!h) return SAVER_ERR_ARG;
  if (!y || !mu || !sf || n < 2 || model < 0 || model > 2 || (mu_len != 1 && mu_len != n) ||
      (sf_len != 1 && sf_len != n))
    return fail(h, SAVER_ERR_ARG, "loglik/calc_var: bad arguments");
  CU(cudaSetDevice(h->device));
  std::vector<double> hm(n), hs(n);
  for (int i = 0; i < n; ++i) {
    hm[i] = mu_len == 1 ? mu[0] : mu[i];
    hs[i] = sf_len == 1 ? sf[0] : sf[i];
  }
  InArr<double> dy, dm, ds;
  CU(dy.bind(y, n, false, h->stream));
  CU(dm.bind(hm.data(), n, false, h->stream));
  CU(ds.bind(hs.data(), n, false, h->stream));
  DevBuf<double> dout, dgrid;
  DevBuf<int> nz;
  CU(dout.alloc(4, h->stream));
  CU(dgrid.alloc(200, h->stream));
  CU(nz.alloc(n, h->stream));
  CU(cudaMemsetAsync(dgrid.p, 0xff, 200 * sizeof(double), h->stream));  // NaN fill
  LogLikArgs A{dy.d, dm.d, ds.d, n, model, mode, param, dout.p, dgrid.p, nz.p};
  CU(launch_loglik(A, h->stream));
  h->launches += 1;
  CU(cudaMemcpyAsync(out4, dout.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (grid200)
    CU(cudaMemcpyAsync(grid200, dgrid.p, 200 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return SAVER_OK;
}

int saver_cuda_loglik(saver_cuda_handle h, int model, double param, const double* y,
                      const double* mu, int mu_len, const double* sf, int sf_len, int n,
                      double* out) {
  if (!out) return fail(h, SAVER_ERR_ARG, "loglik: out is NULL");
  double o4[4];
  int rc = one_gene(h, 0, model, param, y, mu, mu_len, sf, sf_len, n, o4, nullptr);
  if (rc == SAVER_OK) *out = o4[0];
  return rc;
}

int saver_cuda_calc_var(saver_cuda_handle h, int model, const double* y, const double* mu,
                        int mu_len, const double* sf, int sf_len, int n, double* out2,
                        int* flags, double* grid200) {
  if (!out2) return fail(h, SAVER_ERR_ARG, "calc_var: out2 is NULL");
  double o4[4];
  int rc = one_gene(h, 1, model, 0.0, y, mu, mu_len, sf, sf_len, n, o4, grid200);
  if (rc == SAVER_OK) {
    out2[0] = o4[0];
    out2[1] = o4[1];
    if (flags) *flags = (int)o4[2];
  }
  return rc;
}

int saver_cuda_calc_post(saver_cuda_handle h, const double* Y, const double* MU,
                         int mu_is_scalar, const double* sf, int n, int B, double scale_sf,
                         double* est, double* se, double* est_raw, double* se_raw,
                         double* info, double* grid, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!Y || !MU || !sf || !est || n < 2 || B < 0)
    return fail(h, SAVER_ERR_ARG, "calc_post: bad arguments");
  if (B == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  const bool dev = device_ptrs != 0;
  cudaStream_t s = h->stream;
  const size_t nb = (size_t)n * B;
  InArr<double> dY, dMU, dsf;
  CU(dY.bind(Y, nb, dev, s));
  CU(dsf.bind(sf, n, dev, s));
  DevBuf<double> mu_panel;
  const double* mu_d;
  if (mu_is_scalar) {
    CU(dMU.bind(MU, B, dev, s));
    CU(mu_panel.alloc(nb, s));
    fill_panel_kernel<<<592, 256, 0, s>>>(mu_panel.p, dMU.d, n, B);
    CU(cudaGetLastError());
    h->launches += 1;
    mu_d = mu_panel.p;
  } else {
    CU(dMU.bind(MU, nb, dev, s));
    mu_d = dMU.d;
  }
  OutArr<double> oE, oS, oER, oSR, oI, oG;
  CU(oE.bind(est, nb, dev, s));
  CU(oS.bind(se, nb, dev, s));
  CU(oER.bind(est_raw, nb, dev, s));
  CU(oSR.bind(se_raw, nb, dev, s));
  DevBuf<double> info_tmp;
  double* info_d;
  if (info) {
    CU(oI.bind(info, (size_t)kPostInfo * B, dev, s));
    info_d = oI.d;
  } else {
    CU(info_tmp.alloc((size_t)kPostInfo * B, s));
    info_d = info_tmp.p;
  }
  CU(oG.bind(grid, (size_t)600 * B, dev, s));
  if (oG.d) CU(cudaMemsetAsync(oG.d, 0xff, (size_t)600 * B * sizeof(double), s));
  int stage;
  calc_post_smem_bytes(n, &stage);
  DevBuf<int> nz;
  if (stage == 0) CU(nz.alloc(nb, s));
  CalcPostArgs A{dY.d, mu_d, dsf.d, n, B, scale_sf, oE.d, oS.d, oER.d, oSR.d, info_d, oG.d, nz.p};
  CU(launch_calc_post(A, s));
  h->launches += 1;
  CU(oE.fetch(s));
  CU(oS.fetch(s));
  CU(oER.fetch(s));
  CU(oSR.fetch(s));
  CU(oI.fetch(s));
  CU(oG.fetch(s));
  if (!dev) CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

int saver_cuda_set_design(saver_cuda_handle h, const double* xest, int n, int p, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!xest || n < 2 || p < 1) return fail(h, SAVER_ERR_ARG, "set_design: bad arguments");
  CU(cudaSetDevice(h->device));
  InArr<double> dX;
  CU(dX.bind(xest, (size_t)n * p, device_ptrs != 0, h->stream));
  CU(h->design.build(dX.d, n, p, h->stream));
  h->launches += 1;
  CU(cudaStreamSynchronize(h->stream));
  // the other devices of a multi handle: the raw design goes device to device (NVLink where the
  // devices are peers), each device standardises its own copy
  for (saver_cuda_ctx* q : h->peers) {
    const size_t bytes = (size_t)n * p * sizeof(double);
    double* tmp = nullptr;
    if (cudaSetDevice(q->device) != cudaSuccess || cudaMalloc(&tmp, bytes) != cudaSuccess)
      return fail(h, SAVER_ERR_NOMEM, "set_design: no memory for the design on device %d", q->device);
    cudaError_t e = device_ptrs || dX.buf.p
                        ? cudaMemcpyPeer(tmp, q->device, dX.d, h->device, bytes)
                        : cudaMemcpy(tmp, xest, bytes, cudaMemcpyHostToDevice);
    int rc = SAVER_OK;
    if (e != cudaSuccess) rc = fail(h, SAVER_ERR_CUDA, "set_design: copy to device %d: %s", q->device, cudaGetErrorString(e));
    else {
      rc = saver_cuda_set_design(q, tmp, n, p, 1);
      if (rc != SAVER_OK) fail(h, rc, "%s", q->err);
    }
    cudaSetDevice(q->device);
    cudaFree(tmp);
    if (rc != SAVER_OK) return rc;
  }
  return SAVER_OK;
}

int saver_cuda_maxcor(saver_cuda_handle h, const double* Y, int B, const int* self_col,
                      double* out, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!h->design.Xs) return fail(h, SAVER_ERR_STATE, "maxcor: call saver_cuda_set_design first");
  if (!Y || !out || B < 0) return fail(h, SAVER_ERR_ARG, "maxcor: bad arguments");
  if (B == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  const bool dev = device_ptrs != 0;
  cudaStream_t s = h->stream;
  const DesignState& D = h->design;
  InArr<double> dY;
  InArr<int> dSelf;
  OutArr<double> dOut;
  CU(dY.bind(Y, (size_t)D.n * B, dev, s));
  CU(dSelf.bind(self_col, B, dev, s));
  CU(dOut.bind(out, B, dev, s));
  const int chunk = 1024;
  DevBuf<double> R, Cm;
  CU(R.alloc((size_t)D.ldx * chunk, s));
  CU(Cm.alloc((size_t)D.p_pad * chunk, s));
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = B - b0 < chunk ? B - b0 : chunk;
    CU(launch_center_scale(dY.d + (size_t)b0 * D.n, D.n, D.ldx, nb, R.p, s));
    CU(launch_gemm_tn_f64(D.Xs, D.ldx, R.p, D.ldx, Cm.p, D.p_pad, D.p, nb, D.ldx, s));
    CU(launch_colmax(Cm.p, D.p_pad, D.p, nb, dSelf.d ? dSelf.d + b0 : nullptr, dOut.d + b0, s));
    h->launches += 3;
  }
  CU(dOut.fetch(s));
  if (!dev) CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

// ---- expr.predict ----
namespace {

__global__ void pred_mask_kernel(const int* mask, int n, int B, int* fold) {
  const size_t tot = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    fold[i] = mask ? (mask[i % n] != 0) : 1;
}

// glmnet's lambda ratio: alf = max(eps, flmin)^(1/(nlam-1)) with a REAL*4 exponent (Fortran
// fishnet1); evaluated on the host so it matches libm's pow bit for bit.
double lambda_ratio(double flmin, int nlam) {
  const double eqs = flmin > 1e-6 ? flmin : 1e-6;
  return nlam > 1 ? pow(eqs, (double)(1.0f / (float)(nlam - 1))) : 1.0;
}

}  // namespace

int saver_cuda_predict_cv(saver_cuda_handle h, const double* Y, int B, const int* self_col,
                          const int* foldid, int nfolds, int dfmax, int nlambda, double thresh,
                          double* mu, double* fit4, int* status, int* stats, double* cvout,
                          int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!h->design.Xs) return fail(h, SAVER_ERR_STATE, "predict_cv: call saver_cuda_set_design first");
  if (!Y || !foldid || !mu || !fit4 || !status || B < 0 || nfolds < 3 || nfolds > kMaxFolds ||
      nlambda < 1 || nlambda > kNLam || dfmax < 1 || 2 * (long long)dfmax + 20 > kNX || !(thresh > 0))
    return fail(h, SAVER_ERR_ARG, "predict_cv: bad arguments (3 <= nfolds <= %d, 1 <= nlambda <= %d, 1 <= dfmax <= %d, thresh > 0)",
                kMaxFolds, kNLam, (kNX - 20) / 2);
  if (B == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  const bool dev = device_ptrs != 0;
  cudaStream_t s = h->stream;
  const DesignState& D = h->design;
  const size_t n = D.n;
  InArr<double> dY;
  InArr<int> dSelf, dFold;
  OutArr<double> oMu, oFit, oCv;
  OutArr<int> oSt, oStats;
  CU(dY.bind(Y, n * B, dev, s));
  CU(dSelf.bind(self_col, B, dev, s));
  CU(dFold.bind(foldid, n * B, dev, s));
  CU(oMu.bind(mu, n * B, dev, s));
  CU(oFit.bind(fit4, (size_t)4 * B, dev, s));
  CU(oSt.bind(status, B, dev, s));
  CU(oStats.bind(stats, (size_t)4 * B, dev, s));
  CU(oCv.bind(cvout, (size_t)3 * kNLam * B, dev, s));
  LassoParams prm;
  prm.NF = nfolds + 1; prm.nfolds = nfolds; prm.auto_lambda = 1; prm.dfmax = dfmax;
  prm.nlam = nlambda; prm.thr = thresh; prm.maxit = 100000;
  prm.gram = h->opt_gram; prm.strong_slack = h->opt_strong_slack; prm.dbg = h->opt_dbg; prm.inner_scale = h->opt_inner_scale;
  prm.alf_wide = lambda_ratio(0.01, nlambda);
  prm.alf_tall = lambda_ratio(1e-4, nlambda);
  const int chunk = lasso_max_block(prm.NF, D, h->lasso.tune);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    prm.B = B - b0 < chunk ? B - b0 : chunk;
    CU(lasso_run(D, h->lasso, prm, dY.d + n * b0, dFold.d + n * b0, dSelf.d ? dSelf.d + b0 : nullptr,
                 nullptr, nullptr, nullptr, oMu.d + n * b0, oFit.d + (size_t)4 * b0, oSt.d + b0,
                 oStats.d ? oStats.d + (size_t)4 * b0 : nullptr,
                 oCv.d ? oCv.d + (size_t)3 * kNLam * b0 : nullptr, s, &h->launches, &h->gemm_cols));
  }
  CU(oMu.fetch(s)); CU(oFit.fetch(s)); CU(oSt.fetch(s)); CU(oStats.fetch(s)); CU(oCv.fetch(s));
  if (!dev) CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

int saver_cuda_predict_path(saver_cuda_handle h, const double* Y, int B, const int* self_col,
                            const int* pred_mask, const double* lambda_max,
                            const double* lambda_min, int dfmax, double thresh, double* mu,
                            int* status, int* stats, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!h->design.Xs) return fail(h, SAVER_ERR_STATE, "predict_path: call saver_cuda_set_design first");
  if (!Y || !mu || !status || !lambda_max || !lambda_min || B < 0 || dfmax < 1 || 2 * (long long)dfmax + 20 > kNX || !(thresh > 0))
    return fail(h, SAVER_ERR_ARG, "predict_path: bad arguments (1 <= dfmax <= %d, thresh > 0)", (kNX - 20) / 2);
  if (B == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  const bool dev = device_ptrs != 0;
  cudaStream_t s = h->stream;
  const DesignState& D = h->design;
  const size_t n = D.n;
  // lambda.seq = c(exp(seq(log(lmax), log(lmin), by = -0.2)), lmin), sorted decreasing as glmnet
  // does with a user sequence (R/expr_predict.R:62-63); built on the host (always host arrays).
  std::vector<double> lam((size_t)kNLam * B, 0.0);
  std::vector<int> nlam(B, 0);
  for (int g = 0; g < B; ++g) {
    const double lmax = lambda_max[g], lmin = lambda_min[g];
    double* lg = lam.data() + (size_t)kNLam * g;
    int nl = 0;
    if (lmax > 0 && lmin > 0 && lmax == lmax && lmin == lmin) {
      const double from = log(lmax), to = log(lmin);
      int nstep = (int)floor((to - from) / -0.2 + 1e-10);
      if (nstep < 0) nstep = 0;
      if (nstep + 2 > kNLam) return fail(h, SAVER_ERR_ARG, "predict_path: lambda sequence too long");
      for (int i = 0; i <= nstep; ++i) lg[nl++] = exp(from + i * -0.2);
      lg[nl++] = lmin;
      for (int i = 1; i < nl; ++i) {  // insertion sort, decreasing
        const double v = lg[i];
        int j = i - 1;
        while (j >= 0 && lg[j] < v) { lg[j + 1] = lg[j]; --j; }
        lg[j + 1] = v;
      }
    } else {
      lg[nl++] = lmax;  // degenerate input: single value, the solver reports the failure
    }
    nlam[g] = nl;
  }
  InArr<double> dY, dLam, dLmin;
  InArr<int> dSelf, dMask, dNlam;
  OutArr<double> oMu;
  OutArr<int> oSt, oStats;
  CU(dY.bind(Y, n * B, dev, s));
  CU(dSelf.bind(self_col, B, dev, s));
  CU(dMask.bind(pred_mask, n, dev, s));
  CU(dLam.bind(lam.data(), lam.size(), false, s));
  CU(dNlam.bind(nlam.data(), B, false, s));
  CU(dLmin.bind(lambda_min, B, false, s));
  CU(oMu.bind(mu, n * B, dev, s));
  CU(oSt.bind(status, B, dev, s));
  CU(oStats.bind(stats, (size_t)4 * B, dev, s));
  LassoParams prm;
  prm.NF = 1; prm.nfolds = 0; prm.auto_lambda = 0; prm.dfmax = dfmax; prm.nlam = 0;
  prm.thr = thresh; prm.maxit = 100000;
  prm.gram = h->opt_gram; prm.strong_slack = h->opt_strong_slack; prm.dbg = h->opt_dbg; prm.inner_scale = h->opt_inner_scale; prm.alf_wide = prm.alf_tall = 1.0;
  const int chunk = lasso_max_block(prm.NF, D, h->lasso.tune);
  DevBuf<int> fold;
  DevBuf<double> fit4;
  CU(fold.alloc(n * (size_t)(B < chunk ? B : chunk), s));
  CU(fit4.alloc((size_t)4 * B, s));
  pred_mask_kernel<<<592, 256, 0, s>>>(dMask.d, (int)n, B < chunk ? B : chunk, fold.p);
  CU(cudaGetLastError());
  h->launches += 1;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    prm.B = B - b0 < chunk ? B - b0 : chunk;
    CU(lasso_run(D, h->lasso, prm, dY.d + n * b0, fold.p, dSelf.d ? dSelf.d + b0 : nullptr,
                 dLam.d + (size_t)kNLam * b0, dNlam.d + b0, dLmin.d + b0, oMu.d + n * b0,
                 fit4.p + (size_t)4 * b0, oSt.d + b0, oStats.d ? oStats.d + (size_t)4 * b0 : nullptr,
                 nullptr, s, &h->launches, &h->gemm_cols));
  }
  CU(oMu.fetch(s)); CU(oSt.fetch(s)); CU(oStats.fetch(s));
  CU(cudaStreamSynchronize(s));  // host staging vectors die with this frame
  return SAVER_OK;
}

long long saver_cuda_gemm_columns(saver_cuda_handle h) { return h ? h->gemm_cols : 0; }

int saver_cuda_set_option(saver_cuda_handle h, const char* key, double value) {
  if (!h || !key) return SAVER_ERR_ARG;
  if (!strcmp(key, "lasso_gram")) h->opt_gram = value != 0;
  else if (!strcmp(key, "strong_slack")) h->opt_strong_slack = value;
  else if (!strcmp(key, "lasso_dbg")) h->opt_dbg = (int)value;
  else if (!strcmp(key, "inner_scale")) {
    if (!(value > 0 && value <= 1)) return fail(h, SAVER_ERR_ARG, "set_option: inner_scale must be in (0, 1]");
    h->opt_inner_scale = value;
  }
  else return fail(h, SAVER_ERR_ARG, "set_option: unknown key %s", key);
  return SAVER_OK;
}

int saver_cuda_debug_path(saver_cuda_handle h, int g, double* a0, double* dev, int* nin, int* nlp,
                          double* lam) {
  if (!h) return SAVER_ERR_ARG;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  CU(lasso_debug_path(h->lasso, g, a0, dev, nin, nlp, lam));
  return SAVER_OK;
}

int saver_cuda_counters(saver_cuda_handle h, double* out8) {
  if (!h || !out8) return SAVER_ERR_ARG;
  // [8], [9] (optional tail used by scripts/): CTA-lifetime cycles, longest CTA

  out8[0] = h->lasso.ms_round;
  out8[1] = h->lasso.ms_gemm;
  out8[2] = h->lasso.ms_setup;
  out8[3] = h->ms_post;
  out8[4] = (double)h->lasso.cols_streamed;
  out8[5] = (double)h->gemm_cols;
  out8[6] = (double)h->lasso.rounds;
  out8[7] = (double)h->launches;
  return SAVER_OK;
}

int saver_cuda_phase_cycles(saver_cuda_handle h, double* out8) {
  if (!h || !out8) return SAVER_ERR_ARG;
  for (int k = 0; k < 16; ++k) out8[k] = (double)h->lasso.phase_cycles[k];
  return SAVER_OK;
}

int saver_cuda_sweep_hist(saver_cuda_handle h, double* out20) {
  if (!h || !out20) return SAVER_ERR_ARG;
  for (int k = 0; k < 20; ++k) out20[k] = (double)h->lasso.phase_cycles[16 + k];
  return SAVER_OK;
}

int saver_cuda_counters_ext(saver_cuda_handle h, double* out4) {
  if (!h || !out4) return SAVER_ERR_ARG;
  out4[0] = (double)h->lasso.cta_cycles;
  out4[1] = (double)h->lasso.max_cta_cycles;
  out4[2] = (double)h->lasso.gram_fma;
  out4[3] = 0;
  return SAVER_OK;
}

}  // extern "C"
