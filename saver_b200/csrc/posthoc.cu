// posthoc.cu -- consumers of a fitted saver object: sample.saver and cor.genes / cor.cells.
//
// Replaces (reference file:line):
//   R/sample_saver.R:32-79  rgamma(n, est^2/se^2, est/se^2) rounded to 3 decimals, or
//                           rnbinom(n, mu = est, size = est^2/se^2)
//   R/cor_adjust.R:26-66    cor(t(estimate)) (or cor(estimate)) * outer(adj, adj),
//                           adj = sqrt(var / (var + mean(se^2)))
// The reference draws from R's Mersenne-Twister with R's samplers; that stream is not
// reproducible outside R, so parity for sampling is distributional (moments / KS tests).  The
// generator here is counter-based (Philox4x32-10 keyed on the seed, counter = element index), so
// a draw depends only on (seed, rep, gene, cell): results are identical however the genes are
// blocked or sharded over GPUs.
#include <math.h>
#include <stdint.h>

#include "common.cuh"
#include "ctx.h"

using namespace saver;

namespace {

// ------------------------------------------------------------------ Philox4x32-10
// One element (gene, cell, rep) owns the counter stream (idx, rep, block = 0, 1, ...).  Everything is
// kept in registers: the four words of a block are handed out as a uint4 or through predicated
// selects (a dynamically indexed array would live in local memory -- 27 % of the kernel's stall
// samples in the first version, profiles/README.md).
struct Philox {
  uint32_t k0, k1, c0, c1, c2, c3;
  uint4 buf;
  int have;
  __device__ Philox(uint64_t seed, uint64_t idx, uint32_t rep)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0((uint32_t)idx), c1((uint32_t)(idx >> 32)),
        c2(rep), c3(0), have(0) {}
  __device__ __forceinline__ uint4 block() {
    uint32_t a0 = c0, a1 = c1, a2 = c2, a3 = c3, q0 = k0, q1 = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, a0), lo0 = 0xD2511F53u * a0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, a2), lo1 = 0xCD9E8D57u * a2;
      const uint32_t n0 = hi1 ^ a1 ^ q0, n1 = lo1, n2 = hi0 ^ a3 ^ q1, n3 = lo0;
      a0 = n0; a1 = n1; a2 = n2; a3 = n3;
      q0 += 0x9E3779B9u;
      q1 += 0xBB67AE85u;
    }
    c3 += 1;  // next block of this element's private stream
    return make_uint4(a0, a1, a2, a3);
  }
  __device__ __forceinline__ uint32_t next32() {
    if (have == 0) {
      buf = block();
      have = 4;
    }
    const uint32_t w = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
    --have;
    return w;
  }
  // uniform in (0, 1) with 53 random bits
  __device__ __forceinline__ double uniform() {
    const uint64_t a = next32(), b = next32();
    const uint64_t v = ((a << 32) | b) >> 11;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
  }
};

__device__ __forceinline__ double u32_to_unit(uint32_t w) {  // (0, 1), 32 random bits
  return ((double)w + 0.5) * (1.0 / 4294967296.0);
}

// Standard normal from two words: Box-Muller in FP32 (the value is a proposal that is then used in
// FP64 arithmetic; 32 bits of u1 reach 6.7 sigma, and the result of sample.saver is rounded to three
// decimals).  The reference draws with R's own generator, so parity here is distributional anyway.
__device__ __forceinline__ double normal32(uint32_t w0, uint32_t w1) {
  const float u1 = ((float)w0 + 0.5f) * 2.3283064365386963e-10f;   // (0, 1)
  const float th = (float)(w1 >> 8) * (6.283185307179586f / 16777216.0f);
  return (double)(sqrtf(-2.0f * __logf(u1)) * __cosf(th));
}

// Marsaglia-Tsang gamma(shape, 1): one Philox block per proposal (two words -> the normal, one ->
// the acceptance uniform, one -> the shape < 1 boost of the first proposal)
__device__ double gamma_draw(Philox& g, double a) {
  if (!(a > 0.0)) return a == 0.0 ? 0.0 : NAN;
  const bool small = a < 1.0;
  const double a1 = small ? a + 1.0 : a;
  const double d = a1 - 1.0 / 3.0, c = rsqrt(9.0 * d);
  double boost = 1.0;
  for (int it = 0; it < 1000; ++it) {
    const uint4 w = g.block();
    if (it == 0 && small) boost = pow(u32_to_unit(w.w), 1.0 / a);
    const double x = normal32(w.x, w.y);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = u32_to_unit(w.z);
    const double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2 || log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
  }
  return boost * d;
}

// Poisson(lam): multiplication method for small means, Hoermann's PTRS otherwise
__device__ double poisson_draw(Philox& g, double lam) {
  if (!(lam >= 0.0)) return NAN;
  if (lam == 0.0) return 0.0;
  if (lam < 10.0) {
    const double L = exp(-lam);
    double p = 1.0;
    int k = 0;
    do {
      ++k;
      p *= g.uniform();
    } while (p > L);
    return (double)(k - 1);
  }
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int it = 0; it < 1000; ++it) {
    const double u = g.uniform() - 0.5, v = g.uniform();
    const double us = 0.5 - fabs(u);
    const double k = floor((2.0 * a / us + b) * u + lam + 0.43);
    if (us >= 0.07 && v <= vr) return k;
    if (k < 0.0 || (us < 0.013 && v > us)) continue;
    if (log(v) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0))
      return k;
  }
  return floor(lam);
}

__global__ void sample_kernel(const double* __restrict__ est, const double* __restrict__ se,
                              size_t total, size_t elem0, int nb_mode, uint64_t seed, uint32_t rep,
                              double* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const double e = est[i], s = se[i];
    const double s2 = s * s;
    const double shape = e * e / s2;  // NaN for the all-zero genes (0/0), as in R
    Philox g(seed, elem0 + i, rep);
    double v;
    if (nb_mode) {
      // rnbinom(mu, size): Poisson with a Gamma(size, scale = mu / size) mean
      const double lam = gamma_draw(g, shape) * (e / shape);
      v = poisson_draw(g, lam);
    } else {
      // rgamma(shape, rate = est / se^2), rounded to 3 decimals (R/sample_saver.R:57)
      v = gamma_draw(g, shape) * (s2 / e);
      v = rint(v * 1000.0) / 1000.0;
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------ adjusted correlations
// one CTA per vector (gene or cell): standardise to unit norm (so Z^T Z = cor) and compute adj
__global__ void __launch_bounds__(256)
cor_prep_kernel(const double* __restrict__ E, const double* __restrict__ S, int len, int ld_out,
                double* __restrict__ Z, double* __restrict__ adj) {
  __shared__ double scratch[16];
  const int v = blockIdx.x;
  const double* e = E + (size_t)v * len;
  const double* s = S + (size_t)v * len;
  double a[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < len; i += 256) {
    a[0] += e[i];
    a[1] += s[i] * s[i];
  }
  block_sum<2>(a, scratch);
  double mean = a[0] / len, t = 0.0;
  for (int i = threadIdx.x; i < len; i += 256) t += e[i] - mean;
  mean += block_sum1(t, scratch) / len;
  double ss = 0.0;
  for (int i = threadIdx.x; i < len; i += 256) {
    const double d = e[i] - mean;
    ss += d * d;
  }
  ss = block_sum1(ss, scratch);
  const double var = ss / (len - 1), mse = a[1] / len;
  const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : NAN;  // constant vector: cor is NA in R
  double* z = Z + (size_t)v * ld_out;
  for (int i = threadIdx.x; i < ld_out; i += 256) z[i] = i < len ? (e[i] - mean) * inv : 0.0;
  if (threadIdx.x == 0) adj[v] = sqrt(var / (var + mse));
}

// C (M rows x ncols, ld = ldc) *= outer(adj_r, adj_c); C, adj_r, adj_c already point at the block
__global__ void cor_scale_kernel(double* __restrict__ C, int ldc, int M, int ncols,
                                 const double* __restrict__ adj_r, const double* __restrict__ adj_c,
                                 const double* __restrict__ cor_in, int ld_in) {
  const size_t tot = (size_t)M * ncols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i % M), c = (int)(i / M);
    const size_t o = (size_t)c * ldc + r;
    double v = cor_in ? cor_in[(size_t)c * ld_in + r] : C[o];
    if (!cor_in && v == v) v = fmin(1.0, fmax(-1.0, v));  // stats::cor clamps to [-1, 1]; NA stays NA
    C[o] = v * adj_r[r] * adj_c[c];
  }
}

// (rows x cols, column-major, ld = rows) -> transpose
__global__ void transpose_kernel(const double* __restrict__ in, int rows, int cols,
                                 double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = bx + threadIdx.x, c = by + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)c * rows + r];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = by + threadIdx.x, r = bx + j;
    if (r < rows && c < cols) out[(size_t)r * cols + c] = tile[threadIdx.x][j];
  }
}

const int kGrid = 148 * 8;

}  // namespace

extern "C" int saver_cuda_sample(saver_cuda_handle h, const double* est, const double* se, int n, int B,
                                 long long gene0, int nb_mode, unsigned long long seed, int rep_idx,
                                 double* out, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!est || !se || !out || n < 1 || B < 0 || rep_idx < 0 || gene0 < 0)
    return fail(h, SAVER_ERR_ARG, "sample: bad arguments");
  if (B == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  const bool dev = device_ptrs != 0;
  cudaStream_t s = h->stream;
  const size_t nb = (size_t)n * B;
  InArr<double> dE, dS;
  OutArr<double> dO;
  CU(dE.bind(est, nb, dev, s));
  CU(dS.bind(se, nb, dev, s));
  CU(dO.bind(out, nb, dev, s));
  sample_kernel<<<kGrid, 256, 0, s>>>(dE.d, dS.d, nb, (size_t)gene0 * n, nb_mode, seed,
                                     (uint32_t)rep_idx, dO.d);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(dO.fetch(s));
  if (!dev) CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

extern "C" int saver_cuda_cor_adjust(saver_cuda_handle h, const double* est, const double* se, int n,
                                     int G, int by_cells, const double* cor_in, int col0,
                                     int ncols, double* out, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!est || !se || !out || n < 2 || G < 2) return fail(h, SAVER_ERR_ARG, "cor_adjust: bad arguments");
  const int M = by_cells ? n : G;    // vectors being correlated
  const int K = by_cells ? G : n;    // their length
  if (col0 < 0 || ncols < 0 || col0 + ncols > M) return fail(h, SAVER_ERR_ARG, "cor_adjust: bad column block");
  if (ncols == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  const bool dev = device_ptrs != 0;
  cudaStream_t s = h->stream;
  const size_t ng = (size_t)n * G;
  InArr<double> dE, dS, dC;
  OutArr<double> dO;
  CU(dE.bind(est, ng, dev, s));
  CU(dS.bind(se, ng, dev, s));
  CU(dC.bind(cor_in, (size_t)M * ncols, dev, s));
  CU(dO.bind(out, (size_t)M * ncols, dev, s));
  const double *E = dE.d, *S = dS.d;
  DevBuf<double> Et, St;
  if (by_cells) {  // panels are gene-major (n x G); cells need genes contiguous: transpose
    CU(Et.alloc(ng, s));
    CU(St.alloc(ng, s));
    dim3 tg((n + 31) / 32, (G + 31) / 32), tb(32, 8);
    transpose_kernel<<<tg, tb, 0, s>>>(dE.d, n, G, Et.p);
    transpose_kernel<<<tg, tb, 0, s>>>(dS.d, n, G, St.p);
    CU(cudaGetLastError());
    h->launches += 2;
    E = Et.p;
    S = St.p;
  }
  const int ldk = (K + 15) / 16 * 16, M_pad = (M + 127) / 128 * 128;
  DevBuf<double> Z, adj;
  const int Zcols = M_pad + 64;  // the GEMM reads up to round_up(ncols, 64) columns past col0
  CU(Z.alloc((size_t)ldk * Zcols, s));
  CU(adj.alloc(M_pad, s));
  CU(cudaMemsetAsync(Z.p + (size_t)ldk * M, 0, (size_t)ldk * (Zcols - M) * 8, s));
  cor_prep_kernel<<<M, 256, 0, s>>>(E, S, K, ldk, Z.p, adj.p);
  CU(cudaGetLastError());
  h->launches += 1;
  if (!dC.d) {
    CU(launch_gemm_tn_f64(Z.p, ldk, Z.p + (size_t)ldk * col0, ldk, dO.d, M, M, ncols, ldk, s));
    h->launches += 1;
  }
  cor_scale_kernel<<<kGrid, 256, 0, s>>>(dO.d, M, M, ncols, adj.p, adj.p + col0, dC.d, M);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(dO.fetch(s));
  if (!dev) CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

// ---- cor.genes over ranks: the two halves of saver_cuda_cor_adjust as separate calls ----
// saver_cuda_cor_prep standardises this rank's B genes (unit-norm centred rows, so Z^T Z = cor) and
// computes their adj; the ranks all-gather Z / adj (NCCL, outside this library); saver_cuda_cor_block
// then fills one rectangular block of the adjusted correlation from the gathered panel.
extern "C" int saver_cuda_cor_prep(saver_cuda_handle h, const double* est, const double* se, int n,
                                   int B, double* Z, int ldz, double* adj, int device_ptrs) {
  if (!h) return SAVER_ERR_ARG;
  if (!est || !se || !Z || !adj || n < 2 || B < 0 || ldz < n || ldz % 16 != 0)
    return fail(h, SAVER_ERR_ARG, "cor_prep: bad arguments (ldz must be a multiple of 16 >= n)");
  if (B == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  const bool dev = device_ptrs != 0;
  cudaStream_t s = h->stream;
  InArr<double> dE, dS;
  OutArr<double> dZ, dA;
  CU(dE.bind(est, (size_t)n * B, dev, s));
  CU(dS.bind(se, (size_t)n * B, dev, s));
  CU(dZ.bind(Z, (size_t)ldz * B, dev, s));
  CU(dA.bind(adj, (size_t)B, dev, s));
  cor_prep_kernel<<<B, 256, 0, s>>>(dE.d, dS.d, n, ldz, dZ.d, dA.d);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(dZ.fetch(s));
  CU(dA.fetch(s));
  if (!dev) CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

extern "C" int saver_cuda_cor_block(saver_cuda_handle h, const double* Z, int ldz, const double* adj,
                                    int M, int row0, int nrows, int col0, int ncols, double* out,
                                    int ldo) {
  if (!h) return SAVER_ERR_ARG;
  if (!Z || !adj || !out || ldz % 16 != 0 || M < 1 || row0 < 0 || nrows < 0 || row0 + nrows > M ||
      col0 < 0 || ncols < 0 || col0 + ncols > M || ldo < row0 + nrows)
    return fail(h, SAVER_ERR_ARG, "cor_block: bad arguments");
  if (nrows == 0 || ncols == 0) return SAVER_OK;
  CU(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  // device pointers only; Z must be allocated (and zero filled) up to round_up(M, 128) + 64 columns:
  // the GEMM reads whole tiles
  double* C = out + row0;
  CU(launch_gemm_tn_f64(Z + (size_t)ldz * row0, ldz, Z + (size_t)ldz * col0, ldz, C, ldo, nrows, ncols, ldz, s));
  cor_scale_kernel<<<kGrid, 256, 0, s>>>(C, ldo, nrows, ncols, adj + row0, adj + col0, nullptr, 0);
  CU(cudaGetLastError());
  h->launches += 2;
  return SAVER_OK;
}
