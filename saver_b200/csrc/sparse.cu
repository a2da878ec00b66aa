// sparse.cu -- resident sparse counts (R's dgCMatrix: genes x cells, compressed by cell column).
//
// Replaces the Matrix-package calls on the hot path of the reference for sparse input:
//   R/utils.R:11-19    clean.data: x[x < 0.001] <- 0
//   R/utils.R:72       calc.size.factor: colSums(x) / mean(colSums(x))
//   R/saver.R:156-157  rowMeans(x) >= 0.1 ; x.est = t(log((x[good, ] + 1) / sf))   (densified)
//   R/calc_estimate.R:69  as.matrix(x) of a stage's row block
// The matrix is uploaded once; gene blocks are densified on the device straight into the
// gene-major panels the other entry points take.
#include "common.cuh"
#include "ctx.h"

using namespace saver;

namespace {

__global__ void csc_clean_kernel(double* __restrict__ data, size_t nnz) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nnz;
       i += (size_t)gridDim.x * blockDim.x)
    if (data[i] < 0.001) data[i] = 0.0;
}

// one warp per cell column: column sum, and row sums by atomics
__global__ void __launch_bounds__(256)
csc_sums_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                const double* __restrict__ data, int n, double* __restrict__ col_sums,
                double* __restrict__ row_sums) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= n) return;
  double s = 0.0;
  for (int k = indptr[c] + lane; k < indptr[c + 1]; k += 32) {
    const double v = data[k];
    s += v;
    if (v != 0.0) atomicAdd(row_sums + indices[k], v);
  }
  s = warp_sum(s);
  if (lane == 0) col_sums[c] = s;
}

__global__ void fill_map_kernel(int* __restrict__ map, int G) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < G; i += gridDim.x * blockDim.x) map[i] = -1;
}
__global__ void set_map_kernel(int* __restrict__ map, const int* __restrict__ gene_idx, int B) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x)
    map[gene_idx[i]] = i;
}
// background of the panel: 0, or log(1 / sf) where the count is zero
__global__ void panel_fill_kernel(double* __restrict__ out, const double* __restrict__ sf, int n,
                                  int B, int transform) {
  const size_t tot = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x)
    out[i] = transform ? log(1.0 / sf[i % n]) : 0.0;
}
__global__ void __launch_bounds__(256)
csc_scatter_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                   const double* __restrict__ data, const int* __restrict__ map,
                   const double* __restrict__ sf, int n, int transform, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= n) return;
  const double sfc = transform ? sf[c] : 1.0;
  for (int k = indptr[c] + lane; k < indptr[c + 1]; k += 32) {
    const int pos = map[indices[k]];
    if (pos >= 0) {
      const double v = data[k];
      out[(size_t)pos * n + c] = transform ? log((v + 1.0) / sfc) : v;
    }
  }
}

// ---- dense counts (R: a plain matrix) ----
// genes x cells as G contiguous rows of n cells.  The three passes saver() makes over the whole
// matrix on the host -- clean.data's threshold, colSums / rowMeans, and the log-normalised design
// (R/utils.R:11-19,72, R/saver.R:156-157) -- run here instead, so the only full-size transfer is the
// upload of x itself.
template <typename T>
__device__ __forceinline__ double cleaned(T v) {
  const double d = (double)v;
  return d < 0.001 ? 0.0 : d;
}

// one CTA per gene: row sum (fixed reduction order)
template <typename T>
__global__ void __launch_bounds__(256)
dense_row_sums_kernel(const T* __restrict__ x, int n, double* __restrict__ row_sums) {
  __shared__ double scratch[8];
  const T* r = x + (size_t)blockIdx.x * n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += cleaned(r[i]);
  s = block_sum1(s, scratch);
  if (threadIdx.x == 0) row_sums[blockIdx.x] = s;
}

// column sums in two deterministic levels: gene chunk c adds its rows in order into part[c][cell]
template <typename T>
__global__ void __launch_bounds__(256)
dense_col_part_kernel(const T* __restrict__ x, int G, int n, int chunk, double* __restrict__ part) {
  const int cell = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
  if (cell >= n) return;
  const int g0 = c * chunk, g1 = g0 + chunk < G ? g0 + chunk : G;
  double s = 0.0;
  for (int g = g0; g < g1; ++g) s += cleaned(x[(size_t)g * n + cell]);
  part[(size_t)c * n + cell] = s;
}
__global__ void __launch_bounds__(256)
dense_col_final_kernel(const double* __restrict__ part, int nchunk, int n, double* __restrict__ col_sums) {
  const int cell = blockIdx.x * 256 + threadIdx.x;
  if (cell >= n) return;
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += part[(size_t)c * n + cell];
  col_sums[cell] = s;
}

template <typename T>
__global__ void dense_gather_kernel(const T* __restrict__ x, const int* __restrict__ gene_idx, int n,
                                    int B, const double* __restrict__ sf, int transform,
                                    double* __restrict__ out) {
  const size_t tot = (size_t)n * B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < tot;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % n);
    const double v = cleaned(x[(size_t)gene_idx[i / n] * n + c]);
    out[i] = transform ? log((v + 1.0) / sf[c]) : v;
  }
}

}  // namespace

extern "C" int saver_cuda_set_counts_dense(saver_cuda_handle h, const void* x, int dtype, int G, int n) {
  if (!h) return SAVER_ERR_ARG;
  if (!x || G < 1 || n < 1 || (dtype != 0 && dtype != 1))
    return fail(h, SAVER_ERR_ARG, "set_counts_dense: bad arguments (dtype 0 = double, 1 = int32)");
  CU(cudaSetDevice(h->device));
  cudaFree(h->dn_x);
  h->dn_x = nullptr;
  const size_t bytes = (size_t)G * n * (dtype == 0 ? sizeof(double) : sizeof(int));
  CU(cudaMalloc(&h->dn_x, bytes));
  CU(cudaMemcpyAsync(h->dn_x, x, bytes, cudaMemcpyHostToDevice, h->stream));
  h->dn_dtype = dtype;
  h->dn_G = G;
  h->dn_n = n;
  CU(cudaStreamSynchronize(h->stream));
  return SAVER_OK;
}

extern "C" int saver_cuda_dense_sums(saver_cuda_handle h, double* col_sums, double* row_sums) {
  if (!h) return SAVER_ERR_ARG;
  if (!h->dn_x) return fail(h, SAVER_ERR_STATE, "dense_sums: call saver_cuda_set_counts_dense first");
  if (!col_sums || !row_sums) return fail(h, SAVER_ERR_ARG, "dense_sums: null outputs");
  CU(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int G = h->dn_G, n = h->dn_n;
  const int nchunk = G < 64 ? 1 : 64, chunk = (G + nchunk - 1) / nchunk;
  DevBuf<double> dc, dr, part;
  CU(dc.alloc(n, s));
  CU(dr.alloc(G, s));
  CU(part.alloc((size_t)nchunk * n, s));
  const dim3 cg((n + 255) / 256, nchunk);
  if (h->dn_dtype == 0) {
    dense_row_sums_kernel<double><<<G, 256, 0, s>>>((const double*)h->dn_x, n, dr.p);
    dense_col_part_kernel<double><<<cg, 256, 0, s>>>((const double*)h->dn_x, G, n, chunk, part.p);
  } else {
    dense_row_sums_kernel<int><<<G, 256, 0, s>>>((const int*)h->dn_x, n, dr.p);
    dense_col_part_kernel<int><<<cg, 256, 0, s>>>((const int*)h->dn_x, G, n, chunk, part.p);
  }
  dense_col_final_kernel<<<(n + 255) / 256, 256, 0, s>>>(part.p, nchunk, n, dc.p);
  CU(cudaGetLastError());
  h->launches += 3;
  CU(cudaMemcpyAsync(col_sums, dc.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(row_sums, dr.p, (size_t)G * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

extern "C" int saver_cuda_dense_gather(saver_cuda_handle h, const int* gene_idx, int B, const double* sf,
                                       int transform, double* out, int out_on_device) {
  if (!h) return SAVER_ERR_ARG;
  if (!h->dn_x) return fail(h, SAVER_ERR_STATE, "dense_gather: call saver_cuda_set_counts_dense first");
  if (!gene_idx || !out || B < 0 || (transform && !sf))
    return fail(h, SAVER_ERR_ARG, "dense_gather: bad arguments");
  if (B == 0) return SAVER_OK;
  for (int b = 0; b < B; ++b)
    if (gene_idx[b] < 0 || gene_idx[b] >= h->dn_G) return fail(h, SAVER_ERR_ARG, "dense_gather: gene index out of range");
  CU(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int n = h->dn_n;
  InArr<int> dIdx;
  InArr<double> dsf;
  OutArr<double> dO;
  CU(dIdx.bind(gene_idx, B, false, s));
  CU(dsf.bind(sf, n, false, s));
  CU(dO.bind(out, (size_t)n * B, out_on_device != 0, s));
  if (h->dn_dtype == 0)
    dense_gather_kernel<double><<<148 * 8, 256, 0, s>>>((const double*)h->dn_x, dIdx.d, n, B, dsf.d, transform, dO.d);
  else
    dense_gather_kernel<int><<<148 * 8, 256, 0, s>>>((const int*)h->dn_x, dIdx.d, n, B, dsf.d, transform, dO.d);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(dO.fetch(s));
  CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

extern "C" int saver_cuda_release_counts(saver_cuda_handle h) {
  if (!h) return SAVER_ERR_ARG;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  cudaFree(h->dn_x);
  h->dn_x = nullptr;
  h->dn_G = h->dn_n = 0;
  return SAVER_OK;
}

extern "C" int saver_cuda_set_counts_csc(saver_cuda_handle h, const int* indptr, const int* indices,
                                         const double* data, int G, int n) {
  if (!h) return SAVER_ERR_ARG;
  if (!indptr || G < 1 || n < 1) return fail(h, SAVER_ERR_ARG, "set_counts_csc: bad arguments");
  CU(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const size_t nnz = (size_t)indptr[n];
  if (nnz > 0 && (!indices || !data)) return fail(h, SAVER_ERR_ARG, "set_counts_csc: null arrays");
  cudaFree(h->csc_indptr); cudaFree(h->csc_indices); cudaFree(h->csc_data); cudaFree(h->csc_map);
  h->csc_indptr = h->csc_indices = h->csc_map = nullptr;
  h->csc_data = nullptr;
  CU(cudaMalloc(&h->csc_indptr, (size_t)(n + 1) * sizeof(int)));
  CU(cudaMalloc(&h->csc_indices, (nnz ? nnz : 1) * sizeof(int)));
  CU(cudaMalloc(&h->csc_data, (nnz ? nnz : 1) * sizeof(double)));
  CU(cudaMalloc(&h->csc_map, (size_t)G * sizeof(int)));
  CU(cudaMemcpyAsync(h->csc_indptr, indptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
  if (nnz) {
    CU(cudaMemcpyAsync(h->csc_indices, indices, nnz * sizeof(int), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->csc_data, data, nnz * sizeof(double), cudaMemcpyHostToDevice, s));
    csc_clean_kernel<<<148 * 8, 256, 0, s>>>(h->csc_data, nnz);
    CU(cudaGetLastError());
    h->launches += 1;
  }
  h->csc_G = G;
  h->csc_n = n;
  CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

extern "C" int saver_cuda_csc_sums(saver_cuda_handle h, double* col_sums, double* row_sums) {
  if (!h) return SAVER_ERR_ARG;
  if (!h->csc_indptr) return fail(h, SAVER_ERR_STATE, "csc_sums: call saver_cuda_set_counts_csc first");
  if (!col_sums || !row_sums) return fail(h, SAVER_ERR_ARG, "csc_sums: null outputs");
  CU(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  DevBuf<double> dc, dr;
  CU(dc.alloc(h->csc_n, s));
  CU(dr.alloc(h->csc_G, s));
  CU(cudaMemsetAsync(dr.p, 0, (size_t)h->csc_G * sizeof(double), s));
  csc_sums_kernel<<<(h->csc_n + 7) / 8, 256, 0, s>>>(h->csc_indptr, h->csc_indices, h->csc_data,
                                                    h->csc_n, dc.p, dr.p);
  CU(cudaGetLastError());
  h->launches += 1;
  CU(cudaMemcpyAsync(col_sums, dc.p, (size_t)h->csc_n * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(row_sums, dr.p, (size_t)h->csc_G * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}

extern "C" int saver_cuda_csc_gather(saver_cuda_handle h, const int* gene_idx, int B, const double* sf,
                                     int transform, double* out, int out_on_device) {
  if (!h) return SAVER_ERR_ARG;
  if (!h->csc_indptr) return fail(h, SAVER_ERR_STATE, "csc_gather: call saver_cuda_set_counts_csc first");
  if (!gene_idx || !out || B < 0 || (transform && !sf))
    return fail(h, SAVER_ERR_ARG, "csc_gather: bad arguments");
  if (B == 0) return SAVER_OK;
  for (int b = 0; b < B; ++b)
    if (gene_idx[b] < 0 || gene_idx[b] >= h->csc_G) return fail(h, SAVER_ERR_ARG, "csc_gather: gene index out of range");
  CU(cudaSetDevice(h->device));
  cudaStream_t s = h->stream;
  const int n = h->csc_n;
  InArr<int> dIdx;
  InArr<double> dsf;
  OutArr<double> dO;
  CU(dIdx.bind(gene_idx, B, false, s));     // gene_idx and sf are host arrays
  CU(dsf.bind(sf, n, false, s));
  CU(dO.bind(out, (size_t)n * B, out_on_device != 0, s));
  fill_map_kernel<<<148, 256, 0, s>>>(h->csc_map, h->csc_G);
  set_map_kernel<<<(B + 255) / 256, 256, 0, s>>>(h->csc_map, dIdx.d, B);
  panel_fill_kernel<<<148 * 8, 256, 0, s>>>(dO.d, dsf.d, n, B, transform);
  csc_scatter_kernel<<<(n + 7) / 8, 256, 0, s>>>(h->csc_indptr, h->csc_indices, h->csc_data, h->csc_map,
                                                dsf.d, n, transform, dO.d);
  CU(cudaGetLastError());
  h->launches += 4;
  CU(dO.fetch(s));
  CU(cudaStreamSynchronize(s));
  return SAVER_OK;
}
