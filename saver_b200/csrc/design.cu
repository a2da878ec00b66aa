// design.cu -- the shared design matrix and calc.maxcor.
//
// Replaces (reference file:line):
//   R/saver.R:156-157      x.est = t(log((x[good,]+1)/sf)) is handed in by the host; here it is
//                          standardised ONCE over all cells (mean 0, 1/n-variance 1) and kept
//                          resident for every later contraction.  glmnet re-standardises x per
//                          fit; standardisation is affine-invariant, so per-fit moments are
//                          derived from Xs instead (lasso.cu).
//   R/calc_maxcor.R:26-42  max_j |cor(x.est[,j], y)| with the same-named pair zeroed.
#include "common.cuh"
#include "saver_kernels.h"

namespace saver {
namespace {

// one CTA per design column: two-pass mean / sd, write Xs and Xs^2 with zeroed row padding
__global__ void __launch_bounds__(256)
standardize_kernel(const double* __restrict__ X, int n, int ldx, double* __restrict__ Xs,
                   double* __restrict__ Xs2, float* __restrict__ Xf, double* __restrict__ xmean,
                   double* __restrict__ xsd) {
  __shared__ double scratch[8];
  const int j = blockIdx.x;
  const double* x = X + (size_t)j * n;
  double s = 0.0, lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = x[i];
    s += v;
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  s = block_sum1(s, scratch);
  lo = block_min1(lo, scratch);
  hi = block_max1(hi, scratch);
  double mean = s / n;
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += x[i] - mean;
  t = block_sum1(t, scratch);
  mean += t / n;
  double ss = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = x[i] - mean;
    ss += d * d;
  }
  ss = block_sum1(ss, scratch);
  const double sd = (lo == hi) ? 0.0 : sqrt(ss / n);
  const double inv = sd > 0.0 ? 1.0 / sd : 0.0;
  double* o = Xs + (size_t)j * ldx;
  double* o2 = Xs2 + (size_t)j * ldx;
  float* of = Xf + (size_t)j * ldx;
  for (int i = threadIdx.x; i < ldx; i += blockDim.x) {
    const double v = (i < n) ? (x[i] - mean) * inv : 0.0;
    o[i] = v;
    o2[i] = v * v;
    of[i] = (float)v;
  }
  if (threadIdx.x == 0) {
    xmean[j] = mean;
    xsd[j] = sd;
  }
}

// one CTA per gene: yt = (y - mean) / (n * sd_n(y)), so that Xs^T yt = cor(x_j, y)
__global__ void __launch_bounds__(256)
center_scale_kernel(const double* __restrict__ Y, int n, int ldr, double* __restrict__ R) {
  __shared__ double scratch[8];
  const int g = blockIdx.x;
  const double* y = Y + (size_t)g * n;
  double s = 0.0, lo = INFINITY, hi = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = y[i];
    s += v;
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  s = block_sum1(s, scratch);
  lo = block_min1(lo, scratch);
  hi = block_max1(hi, scratch);
  double mean = s / n, t = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += y[i] - mean;
  mean += block_sum1(t, scratch) / n;
  double ss = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = y[i] - mean;
    ss += d * d;
  }
  ss = block_sum1(ss, scratch);
  const double inv = (lo == hi || ss <= 0.0) ? 0.0 : 1.0 / (sqrt(ss / n) * n);  // sd(y)==0 -> cor NA -> 0
  double* r = R + (size_t)g * ldr;
  for (int i = threadIdx.x; i < ldr; i += blockDim.x) r[i] = (i < n) ? (y[i] - mean) * inv : 0.0;
}

// one CTA per gene: max_j |C[j, g]| skipping the gene's own column
__global__ void __launch_bounds__(256)
colmax_kernel(const double* __restrict__ C, int ldc, int p, const int* __restrict__ self,
              double* __restrict__ out) {
  __shared__ double scratch[8];
  const int g = blockIdx.x;
  const int sc = self ? self[g] : -1;
  const double* c = C + (size_t)g * ldc;
  double m = 0.0;
  for (int j = threadIdx.x; j < p; j += blockDim.x) {
    double v = fabs(c[j]);
    if (j == sc || isnan(v)) v = 0.0;
    m = fmax(m, fmin(v, 1.0));
  }
  m = block_max1(m, scratch);
  if (threadIdx.x == 0) out[g] = m;
}

}  // namespace

cudaError_t DesignState::build(const double* X_dev, int n_, int p_, cudaStream_t s) {
  release();
  n = n_;
  p = p_;
  ldx = (n + 15) / 16 * 16;
  p_pad = (p + 127) / 128 * 128;
  const size_t bytes = (size_t)ldx * p_pad * sizeof(double);
  cudaError_t e;
  if ((e = cudaMalloc(&Xs, bytes)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&Xs2, bytes)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&Xf, bytes / 2)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(Xf, 0, bytes / 2, s)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&xmean, p_pad * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&xsd, p_pad * sizeof(double))) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(Xs, 0, bytes, s)) != cudaSuccess) return e;
  if ((e = cudaMemsetAsync(Xs2, 0, bytes, s)) != cudaSuccess) return e;
  standardize_kernel<<<p, 256, 0, s>>>(X_dev, n, ldx, Xs, Xs2, Xf, xmean, xsd);
  return cudaGetLastError();
}

cudaError_t launch_center_scale(const double* Y, int n, int ldr, int B, double* R, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  center_scale_kernel<<<B, 256, 0, s>>>(Y, n, ldr, R);
  return cudaGetLastError();
}

cudaError_t launch_colmax(const double* C, int ldc, int p, int B, const int* self, double* out,
                          cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  colmax_kernel<<<B, 256, 0, s>>>(C, ldc, p, self, out);
  return cudaGetLastError();
}

}  // namespace saver
