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

}  // namespace

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
