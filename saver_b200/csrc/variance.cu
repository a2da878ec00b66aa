// variance.cu -- per-gene gamma-prior variance fit + Gamma posterior, one CTA per gene.
//
// Replaces (reference file:line, /root/reference):
//   R/calc_loglik.R:26-75         calc.loglik.a/b/k   -> LogLik::eval
//   R/optimize_variance.R:24-134  calc.a/b/k          -> fit_model (Brent fmin, zeroin, grid)
//   R/calc_posterior.R:26-68      calc.post           -> calc_post_kernel
// and the base-R routines those call: stats::optimize (Brent_fmin), stats::uniroot
// (R_zeroin2), both with tol = DBL_EPSILON^0.25.
//
// Layout: panels are n x B column-major (one gene = one contiguous column of n cells),
// fp64.  One CTA owns one gene; its y / mu / sf columns are staged in shared memory
// (STAGE 2), or mu+sf only (STAGE 1, y re-read through L1/L2), or nothing (STAGE 0) when
// n outgrows 227 KB.  Every log-likelihood evaluation is one strided pass over the staged
// cells with a warp-shuffle + shared-memory tree reduction; the scalar optimiser state is
// replicated in every thread (all threads see bit-identical reductions, so control flow is
// uniform).  Cells with y == 0 contribute lgamma(alpha) - lgamma(alpha) == 0 to models b
// and k, so the lgamma pair runs only over a compacted list of non-zero cells.
//
// Algorithmic HBM bytes per gene: read y, mu (16 n) + write est, se (16 n) = 32 n
// (24 n with estimates.only); sf is shared by all genes.
#include "common.cuh"
#include "saver_kernels.h"

namespace saver {

namespace {

// CTA size: 256 threads, 512 when a gene has more than 4,096 cells (one CTA per SM there -- the staged
// columns fill shared memory -- so the extra warps are what hides the FP64 dependency chains of log).
constexpr int kThreads = 256;      // loglik_kernel and small n
constexpr int kThreadsBig = 512;
constexpr int kMaxNW = kThreadsBig / 32;
constexpr double kBrentTol = 1.220703125e-4;  // .Machine$double.eps^0.25

struct Gene {
  const double* y;   // raw counts
  const double* mu;  // prediction
  const double* sf;  // size factors
  const int* nz;     // indices of cells with y != 0
  int n, nnz;
  double sum_logmu;     // sum log(mu)
  double sum_mu2logmu;  // sum mu^2 log(mu)
  double* scratch;      // block reduction scratch
};

// lgamma(y + al) - lgamma(al) for a cell with y != 0.  For the integer counts the path sees this is
// log prod_{j < y} (al + j): one log instead of two lgamma calls, and without the cancellation of
// two large lgamma values (the reference's own arithmetic loses ~1e-7 absolute there when al is
// large).  Non-integer y, y > 32 or a product that could overflow fall back to the lgamma pair.
__device__ __forceinline__ double lgamma_diff(double y, double al) {
  if (y <= 32.0 && y == floor(y) && al < 1.0e9) {
    double pr = al;
    for (double j = 1.0; j < y; j += 1.0) pr *= al + j;
    return log(pr);
  }
  return lgamma(y + al) - lgamma(al);
}

// -calc.loglik.{a,b,k}(param): identical on every thread.
__device__ double loglik(int model, double prm, const Gene& g) {
  const int n = g.n;
  if (model == 0) {
    // R/calc_loglik.R:34-39
    const double ia = 1.0 / prm;
    double acc[2] = {0.0, 0.0};  // sum [lgamma(y+1/a) - lgamma(1/a)] over y!=0 ; sum (y+1/a) log(sf + 1/(a mu))
    for (int t = threadIdx.x; t < g.nnz; t += blockDim.x) acc[0] += lgamma_diff(g.y[g.nz[t]], ia);
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      acc[1] += (g.y[i] + ia) * log(g.sf[i] + 1.0 / (prm * g.mu[i]));
    block_sum<2>(acc, g.scratch);
    const double f1 = n / prm * log(ia);
    const double f2 = -(ia * g.sum_logmu);
    // f3 + f4 of R/calc_loglik.R:36-37 = -n lgamma(1/a) + sum lgamma(y + 1/a): the cells with y == 0
    // cancel exactly, the others are the differences accumulated above
    const double f34 = acc[0];
    const double f5 = -acc[1];
    return -(f1 + f2 + f34 + f5);
  }
  if (model == 1) {
    // R/calc_loglik.R:52-56
    const double ib = 1.0 / prm, lib = log(ib);
    double acc[3] = {0.0, 0.0, 0.0};  // sum mu/b log(1/b); sum[lgamma(y+mu/b)-lgamma(mu/b)]; sum (y+mu/b) log(sf+1/b)
    for (int t = threadIdx.x; t < g.nnz; t += blockDim.x) {
      const int i = g.nz[t];
      const double mb = g.mu[i] / prm;
      acc[1] += lgamma_diff(g.y[i], mb);
    }
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) {  // unrolled: the logs of four cells overlap
      const double mb = g.mu[i] / prm;
      acc[0] += mb * lib;
      acc[2] += (g.y[i] + mb) * log(g.sf[i] + ib);
    }
    block_sum<3>(acc, g.scratch);
    return -(acc[0] + acc[1] - acc[2]);
  }
  // R/calc_loglik.R:69-74
  const double lk = log(prm);
  double acc[3] = {0.0, 0.0, 0.0};  // sum mu^2 log(k)/k; lgamma pair; sum (y+mu^2/k) log(sf+mu/k)
  for (int t = threadIdx.x; t < g.nnz; t += blockDim.x) {
    const int i = g.nz[t];
    const double m = g.mu[i], mk = m * m / prm;
    acc[1] += lgamma_diff(g.y[i], mk);
  }
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = g.mu[i], m2 = m * m;
    acc[0] += m2 * lk / prm;
    acc[2] += (g.y[i] + m2 / prm) * log(g.sf[i] + m / prm);
  }
  block_sum<3>(acc, g.scratch);
  return -(g.sum_mu2logmu / prm - acc[0] + acc[1] - acc[2]);
}

// The 100-point grid of the fallback branch (R/optimize_variance.R:46-52) as ONE pass: the
// parameters are known up front, so warp w takes grid points w, w + 8, ... (13 per warp, their
// accumulators in registers) and streams over all cells; no block barrier until the end.  Per point
// the sums run over the same cells with the same per-cell expressions as loglik(), in a different
// association order (per-lane strided partial sums), i.e. they agree with it to rounding.
// out[i] = -nll(x_i) for i < 100 (every thread reads them from shared memory afterwards).
constexpr int kGridPts = 100;
constexpr int kGridPerWarp = (kGridPts + 8 - 1) / 8;  // 8 warps at least

__device__ void loglik_grid(int model, const double* xs, const Gene& g, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = g.n, kNW = blockDim.x >> 5;
  const int npw = warp < kGridPts ? (kGridPts - warp + kNW - 1) / kNW : 0;  // grid points of this warp
  double prm[kGridPerWarp], c1[kGridPerWarp], acc[kGridPerWarp];
#pragma unroll
  for (int q = 0; q < kGridPerWarp; ++q) {
    const int i = warp + q * kNW;
    prm[q] = xs[i < kGridPts ? i : kGridPts - 1];
    c1[q] = model == 0 ? 1.0 / prm[q] : model == 1 ? 1.0 / prm[q] : prm[q];
    acc[q] = 0.0;
  }
  for (int t = lane; t < g.nnz; t += 32) {
    const int i = g.nz[t];
    const double y = g.y[i], m = g.mu[i];
#pragma unroll
    for (int q = 0; q < kGridPerWarp; ++q)
      if (q < npw) {
        const double al = model == 0 ? c1[q] : model == 1 ? m / prm[q] : m * m / prm[q];
        acc[q] += lgamma_diff(y, al);
      }
  }
  if (model == 0) {
    for (int i = lane; i < n; i += 32) {
      const double y = g.y[i], m = g.mu[i], sfi = g.sf[i];
#pragma unroll
      for (int q = 0; q < kGridPerWarp; ++q)
        if (q < npw) acc[q] -= (y + c1[q]) * log(sfi + 1.0 / (prm[q] * m));
    }
  } else if (model == 1) {
    double lib[kGridPerWarp];
#pragma unroll
    for (int q = 0; q < kGridPerWarp; ++q) lib[q] = log(c1[q]);
    for (int i = lane; i < n; i += 32) {
      const double y = g.y[i], m = g.mu[i], sfi = g.sf[i];
#pragma unroll
      for (int q = 0; q < kGridPerWarp; ++q)
        if (q < npw) {
          const double mb = m / prm[q];
          acc[q] += mb * lib[q] - (y + mb) * log(sfi + c1[q]);
        }
    }
  } else {
    double lk[kGridPerWarp];
#pragma unroll
    for (int q = 0; q < kGridPerWarp; ++q) lk[q] = log(prm[q]);
    for (int i = lane; i < n; i += 32) {
      const double y = g.y[i], m = g.mu[i], sfi = g.sf[i], m2 = m * m;
#pragma unroll
      for (int q = 0; q < kGridPerWarp; ++q)
        if (q < npw) acc[q] -= m2 * lk[q] / prm[q] + (y + m2 / prm[q]) * log(sfi + m / prm[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < kGridPerWarp; ++q) {
    double v = warp_sum(acc[q]);
    const int i = warp + q * kNW;
    if (model == 0) v += n / prm[q] * log(c1[q]) - c1[q] * g.sum_logmu;
    else if (model == 2) v += g.sum_mu2logmu / prm[q];
    if (lane == 0 && i < kGridPts) out[i] = v;
  }
  __syncthreads();
}

__device__ __forceinline__ double guard(double v) { return isfinite(v) ? v : DBL_MAX; }

// stats::optimize(f, c(lo, hi)) -> Brent_fmin.
__device__ double brent_fmin(int model, double lo, double hi, const Gene& g, int& nev) {
  const double gold = (3.0 - sqrt(5.0)) * 0.5;
  const double sqeps = sqrt(DBL_EPSILON);
  const double tol3 = kBrentTol / 3.0;
  double a = lo, b = hi;
  double x = a + gold * (b - a), w = x, v = x;
  double step = 0.0, prev = 0.0;
  double fx = guard(loglik(model, x, g)), fw = fx, fv = fx;
  ++nev;
  for (int it = 0; it < 1000; ++it) {
    const double mid = (a + b) * 0.5;
    const double tol1 = sqeps * fabs(x) + tol3;
    const double tol2 = tol1 * 2.0;
    if (fabs(x - mid) <= tol2 - (b - a) * 0.5) break;
    double p = 0.0, q = 0.0, r = 0.0;
    if (fabs(prev) > tol1) {
      r = (x - w) * (fx - fv);
      q = (x - v) * (fx - fw);
      p = (x - v) * q - (x - w) * r;
      q = (q - r) * 2.0;
      if (q > 0.0) p = -p; else q = -q;
      r = prev;
      prev = step;
    }
    double u;
    if (fabs(p) >= fabs(q * 0.5 * r) || p <= q * (a - x) || p >= q * (b - x)) {
      prev = (x < mid) ? b - x : a - x;
      step = gold * prev;
    } else {
      step = p / q;
      u = x + step;
      if (u - a < tol2 || b - u < tol2) {
        step = tol1;
        if (x >= mid) step = -step;
      }
    }
    if (fabs(step) >= tol1) u = x + step;
    else if (step > 0.0) u = x + tol1;
    else u = x - tol1;
    const double fu = guard(loglik(model, u, g));
    ++nev;
    if (fu <= fx) {
      if (u < x) b = x; else a = x;
      v = w; fv = fw;
      w = x; fw = fx;
      x = u; fx = fu;
    } else {
      if (u < x) a = u; else b = u;
      if (fu <= fw || w == x) {
        v = w; fv = fw;
        w = u; fw = fu;
      } else if (fu <= fv || v == x || v == w) {
        v = u; fv = fu;
      }
    }
  }
  return x;
}

// stats::uniroot -> R_zeroin2 on f(x) = nll(x) + shift.
__device__ double zeroin(int model, double ax, double bx, double fa, double fb, double shift,
                         const Gene& g, int& nev) {
  double a = ax, b = bx, c = a, fc = fa;
  if (fa == 0.0) return a;
  if (fb == 0.0) return b;
  for (int it = 0; it <= 1000; ++it) {
    const double prev_step = b - a;
    if (fabs(fc) < fabs(fb)) {
      a = b; b = c; c = a;
      fa = fb; fb = fc; fc = fa;
    }
    const double tol_act = 2 * DBL_EPSILON * fabs(b) + kBrentTol / 2;
    double new_step = (c - b) / 2;
    if (fabs(new_step) <= tol_act || fb == 0.0) return b;
    if (fabs(prev_step) >= tol_act && fabs(fa) > fabs(fb)) {
      double p, q;
      const double cb = c - b;
      if (a == c) {
        const double t1 = fb / fa;
        p = cb * t1;
        q = 1.0 - t1;
      } else {
        const double qq = fa / fc, t1 = fb / fc, t2 = fb / fa;
        p = t2 * (cb * qq * (qq - t1) - (b - a) * (t1 - 1.0));
        q = (qq - 1.0) * (t1 - 1.0) * (t2 - 1.0);
      }
      if (p > 0.0) q = -q; else p = -p;
      if (p < (0.75 * cb * q - fabs(tol_act * q) / 2) && p < fabs(prev_step * q / 2)) new_step = p / q;
    }
    if (fabs(new_step) < tol_act) new_step = (new_step > 0.0) ? tol_act : -tol_act;
    a = b; fa = fb;
    b += new_step;
    fb = loglik(model, b, g) + shift;
    ++nev;
    if ((fb > 0 && fc > 0) || (fb < 0 && fc < 0)) {
      c = a; fc = fa;
    }
  }
  return b;
}

struct FitResult {
  double param;  // 1/a, 1/b or k
  double nll;
  int flags;     // bit0 grid branch, bit1 uniroot, bit2 error (R would have raised)
  int nev;
};

// calc.a / calc.b / calc.k (R/optimize_variance.R).  upper = search interval end.
// grid_out (optional): 200 doubles, grid x then grid log-lik.
__device__ FitResult fit_model(int model, double upper, const Gene& g, double* grid_s,
                               double* grid_out) {
  FitResult r;
  r.flags = 0;
  r.nev = 0;
  double par = brent_fmin(model, 0.0, upper, g, r.nev);
  double nll = loglik(model, par, g);
  const double ll_min = -loglik(model, 1e-05, g);
  const double ll_mle = -nll;  // the reference re-evaluates at the same point
  const double ll_max = -loglik(model, upper, g);
  r.nev += 3;
  bool err = false;
  if (isnan(ll_mle - ll_min)) err = true;
  else if (ll_mle - ll_min < 0.5) {
    r.flags |= 1;
    double gmax = upper;
    if (isnan(ll_mle - ll_max)) err = true;
    else if (ll_mle - ll_max > 10) {
      const double shift = ll_mle - 10;
      const double flo = -ll_min + shift, fhi = -ll_max + shift;
      if (isnan(flo) || isnan(fhi) || flo * fhi > 0) err = true;
      else {
        gmax = zeroin(model, 1e-05, upper, flo, fhi, shift, g, r.nev);
        r.flags |= 2;
      }
    }
    if (!err) {
      // samp = (exp((exp(ppoints(100)) - 1)/2) - 1) * gmax ; weights exp(ll - min ll)
      double llmin = INFINITY;
      bool bad = false;
      __syncthreads();  // grid_s may still be read by the previous model's epilogue
      if (threadIdx.x < 100) {
        const double pp = (threadIdx.x + 1 - 0.5) / 100.0;
        grid_s[threadIdx.x] = (exp((exp(pp) - 1) / 2) - 1) * gmax;
      }
      __syncthreads();
      loglik_grid(model, grid_s, g, grid_s + 100);
      for (int i = 0; i < 100; ++i) {
        const double ll = grid_s[100 + i];
        if (ll < llmin) llmin = ll;
        if (isnan(ll)) bad = true;
      }
      r.nev += 100;
      if (grid_out && threadIdx.x < 200) grid_out[threadIdx.x] = grid_s[threadIdx.x];
      if (bad) err = true;
      else {
        double tot = 0, acc = 0;
        for (int i = 0; i < 100; ++i) tot += exp(grid_s[100 + i] - llmin);
        for (int i = 0; i < 100; ++i) acc += grid_s[i] * (exp(grid_s[100 + i] - llmin) / tot);
        if (!isfinite(acc)) err = true;
        else {
          par = acc;  // expectation of the reference's resampled mean (optimize_variance.R:53)
          nll = loglik(model, par, g);
          ++r.nev;
        }
      }
      __syncthreads();
    }
  }
  if (err) {
    r.flags |= 4;
    r.param = 0.0;
    r.nll = INFINITY;
  } else {
    r.param = (model == 2) ? par : 1.0 / par;
    r.nll = nll;
  }
  return r;
}

template <int STAGE>
__global__ void __launch_bounds__(kThreadsBig)
calc_post_kernel(CalcPostArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[4 * kMaxNW];
  __shared__ double grid_s[200];
  __shared__ int warp_cnt[kMaxNW + 1];

  const int gi = blockIdx.x;
  const int n = A.n;
  const double* gy = A.Y + (size_t)gi * n;
  const double* gmu = A.MU + (size_t)gi * n;

  // ---- carve shared memory ----
  double* sm_d = reinterpret_cast<double*>(smem_raw);
  double *sy = nullptr, *smu = nullptr, *ssf = nullptr;
  int* snz;
  if (STAGE == 2) {
    sy = sm_d; smu = sm_d + n; ssf = sm_d + 2 * n;
    snz = reinterpret_cast<int*>(sm_d + 3 * n);
  } else if (STAGE == 1) {
    smu = sm_d; ssf = sm_d + n;
    snz = reinterpret_cast<int*>(sm_d + 2 * n);
  } else {
    snz = A.nz_ws + (size_t)gi * n;
  }

  // ---- stage + first-pass statistics ----
  double st[4] = {0.0, 0.0, 0.0, 0.0};  // sum y, sum y/sf, sum mu, (unused)
  double mu_lo = INFINITY, mu_hi = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double y = gy[i], m = gmu[i], s = A.sf[i];
    if (STAGE == 2) sy[i] = y;
    if (STAGE >= 1) { smu[i] = m; ssf[i] = s; }
    st[0] += y;
    st[1] += y / s;
    st[2] += m;
    mu_lo = fmin(mu_lo, m);
    mu_hi = fmax(mu_hi, m);
  }
  block_sum<4>(st, scratch);
  mu_lo = block_min1(mu_lo, scratch);
  mu_hi = block_max1(mu_hi, scratch);

  double* est = A.EST + (size_t)gi * n;
  double* se = A.SE ? A.SE + (size_t)gi * n : nullptr;
  double* info = A.INFO + (size_t)gi * kPostInfo;

  if (st[0] == 0.0) {  // R/calc_posterior.R:34-36
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      est[i] = 0.0;
      if (se) se[i] = 0.0;
    }
    if (threadIdx.x == 0) {
      info[0] = -1; info[1] = 0; info[2] = info[3] = info[4] = NAN; info[5] = 0; info[6] = 0; info[7] = 0;
    }
    return;
  }

  Gene g;
  g.y = (STAGE == 2) ? sy : gy;
  g.mu = (STAGE >= 1) ? smu : gmu;
  g.sf = (STAGE >= 1) ? ssf : A.sf;
  g.n = n;
  g.scratch = scratch;

  // ---- deterministic compaction of the non-zero cells (ascending index order) ----
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;
    for (int start = 0; start < n; start += blockDim.x) {
      const int i = start + threadIdx.x;
      const bool nzf = (i < n) && (g.y[i] != 0.0);
      const unsigned bal = __ballot_sync(0xffffffffu, nzf);
      if (lane == 0) warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int off = base;
      for (int w = 0; w < warp; ++w) off += warp_cnt[w];
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_cnt[w];
      if (nzf) snz[off + __popc(bal & ((1u << lane) - 1))] = i;
      base += tot;
      __syncthreads();
    }
    g.nz = snz;
    g.nnz = base;
  }

  // ---- second-pass statistics: var(y/sf), sum log mu, sum mu^2 log mu ----
  const double mean_ysf = st[1] / n;
  double s2[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = g.y[i] / g.sf[i] - mean_ysf;
    const double m = g.mu[i], lm = log(m);
    s2[0] += d * d;
    s2[1] += lm;
    s2[2] += m * m * lm;
  }
  block_sum<3>(s2, scratch);
  const double var_ysf = s2[0] / (n - 1);
  g.sum_logmu = s2[1];
  g.sum_mu2logmu = s2[2];

  const double up_a = var_ysf / (mean_ysf * mean_ysf);
  const double up_b = var_ysf / mean_ysf;
  const double up_k = var_ysf;

  double* grid_out = A.GRID ? A.GRID + (size_t)gi * 600 : nullptr;
  int method, fbmask = 0, nev = 0, errmask = 0;
  double param, nll3[3] = {NAN, NAN, NAN};
  if (mu_lo == mu_hi) {  // var(mu) == 0, R/calc_posterior.R:37-39
    FitResult rb;
    if (st[2] == 0.0) { rb.param = 0; rb.nll = 0; rb.flags = 0; rb.nev = 0; }
    else rb = fit_model(1, up_b, g, grid_s, grid_out ? grid_out + 200 : nullptr);
    method = 3;
    param = rb.param;
    nll3[1] = rb.nll;
    fbmask = (rb.flags & 1) << 1;
    errmask = (rb.flags & 4) ? 2 : 0;
    nev = rb.nev;
  } else {
    FitResult r[3];
    r[0] = fit_model(0, up_a, g, grid_s, grid_out);
    if (st[2] == 0.0) { r[1].param = 0; r[1].nll = 0; r[1].flags = 0; r[1].nev = 0; }
    else r[1] = fit_model(1, up_b, g, grid_s, grid_out ? grid_out + 200 : nullptr);
    r[2] = fit_model(2, up_k, g, grid_s, grid_out ? grid_out + 400 : nullptr);
    method = -1;
    for (int m = 0; m < 3; ++m) {
      nll3[m] = r[m].nll;
      fbmask |= (r[m].flags & 1) << m;
      if (r[m].flags & 4) errmask |= 1 << m;
      nev += r[m].nev;
      if (!isnan(r[m].nll) && (method < 0 || r[m].nll < r[method].nll)) method = m;
    }
    if (method < 0) method = 0;
    param = r[method].param;
  }

  // ---- Gamma posterior (R/calc_posterior.R:62-67), coalesced stores ----
  const double scale = 1000.0 * A.scale_sf;
  double* est_raw = A.EST_RAW ? A.EST_RAW + (size_t)gi * n : nullptr;
  double* se_raw = A.SE_RAW ? A.SE_RAW + (size_t)gi * n : nullptr;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = g.mu[i];
    double al, be;
    if (method == 0) { al = param; be = param / m; }
    else if (method == 2) { al = m * m / param; be = m / param; }
    else { be = param; al = m * param; }
    const double pa = al + g.y[i], pb = be + g.sf[i];
    const double lam = pa / pb, sd = sqrt(pa / (pb * pb));
    est[i] = ceil(lam * scale) / 1000.0;
    if (se) se[i] = ceil(sd * scale) / 1000.0;
    if (est_raw) est_raw[i] = lam;
    if (se_raw) se_raw[i] = sd;
  }
  if (threadIdx.x == 0) {
    info[0] = method; info[1] = param;
    info[2] = nll3[0]; info[3] = nll3[1]; info[4] = nll3[2];
    info[5] = fbmask; info[6] = nev; info[7] = errmask;
  }
}

template <int STAGE>
__global__ void __launch_bounds__(kThreads) loglik_kernel(LogLikArgs A) {
  // One CTA: calc.loglik.* (mode 0) or calc.a/b/k (mode 1) for a single gene.
  __shared__ double scratch[4 * kMaxNW];
  __shared__ double grid_s[200];
  __shared__ int cnt;
  const int n = A.n;
  Gene g;
  g.y = A.y; g.mu = A.mu; g.sf = A.sf; g.n = n; g.scratch = scratch;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  // single-gene utility path: serial compaction by thread 0 keeps ascending order
  if (threadIdx.x == 0) {
    int c = 0;
    for (int i = 0; i < n; ++i) if (A.y[i] != 0.0) A.nz_ws[c++] = i;
    cnt = c;
  }
  __syncthreads();
  g.nz = A.nz_ws;
  g.nnz = cnt;
  double st[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double m = A.mu[i], lm = log(m);
    st[0] += A.y[i] / A.sf[i];
    st[1] += lm;
    st[2] += m * m * lm;
    st[3] += m;
  }
  block_sum<4>(st, scratch);
  g.sum_logmu = st[1];
  g.sum_mu2logmu = st[2];
  if (A.mode == 0) {
    const double v = loglik(A.model, A.param, g);
    if (threadIdx.x == 0) A.out[0] = v;
    return;
  }
  const double mean_ysf = st[0] / n;
  double ss = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = A.y[i] / A.sf[i] - mean_ysf;
    ss += d * d;
  }
  ss = block_sum1(ss, scratch);
  const double var_ysf = ss / (n - 1);
  const double upper = A.model == 0 ? var_ysf / (mean_ysf * mean_ysf)
                     : A.model == 1 ? var_ysf / mean_ysf : var_ysf;
  FitResult r;
  if (A.model == 1 && st[3] == 0.0) { r.param = 0; r.nll = 0; r.flags = 0; r.nev = 0; }
  else r = fit_model(A.model, upper, g, grid_s, A.grid);
  if (threadIdx.x == 0) {
    A.out[0] = r.param; A.out[1] = r.nll; A.out[2] = r.flags; A.out[3] = r.nev;
  }
}

}  // namespace

size_t calc_post_smem_bytes(int n, int* stage) {
  const size_t full = (size_t)n * (3 * sizeof(double) + sizeof(int));
  const size_t part = (size_t)n * (2 * sizeof(double) + sizeof(int));
  const size_t cap = 200 * 1024;
  if (full <= cap) { *stage = 2; return full; }
  if (part <= cap) { *stage = 1; return part; }
  *stage = 0;
  return 0;
}

cudaError_t launch_calc_post(const CalcPostArgs& A, cudaStream_t stream) {
  int stage;
  const size_t smem = calc_post_smem_bytes(A.n, &stage);
  if (A.B <= 0) return cudaSuccess;
  const int tpb = A.n > 4096 ? kThreadsBig : kThreads;
  cudaError_t e;
  if (stage == 2) {
    e = cudaFuncSetAttribute(calc_post_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    calc_post_kernel<2><<<A.B, tpb, smem, stream>>>(A);
  } else if (stage == 1) {
    e = cudaFuncSetAttribute(calc_post_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    calc_post_kernel<1><<<A.B, tpb, smem, stream>>>(A);
  } else {
    if (!A.nz_ws) return cudaErrorInvalidValue;
    calc_post_kernel<0><<<A.B, tpb, 0, stream>>>(A);
  }
  return cudaGetLastError();
}

cudaError_t launch_loglik(const LogLikArgs& A, cudaStream_t stream) {
  loglik_kernel<0><<<1, kThreads, 0, stream>>>(A);
  return cudaGetLastError();
}

}  // namespace saver
