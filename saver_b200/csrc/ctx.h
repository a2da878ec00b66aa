// ctx.h -- the handle behind include/saver_cuda.h and the marshalling helpers shared by the
// C-ABI translation units (saver_cuda.cu, estimate.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <vector>

#include "../../include/saver_cuda.h"
#include "saver_kernels.h"

struct saver_cuda_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  long long gemm_cols = 0;  // problem columns pushed through the gradient GEMM (lasso rounds)
  double ms_post = 0, ms_maxcor = 0;
  int opt_gram = 1;            // saver_cuda_set_option("lasso_gram")
  double opt_strong_slack = 1.0;
  double opt_inner_scale = 1.0;  // saver_cuda_set_option("inner_scale")
  int opt_dbg = 0;  // device time of the variance/posterior and maxcor phases
  char err[512] = {0};
  // further devices behind this handle (saver_cuda_create_multi); empty for a one-device handle
  std::vector<saver_cuda_ctx*> peers;
  saver::DesignState design;
  saver::LassoWs lasso;
  // resident sparse counts (sparse.cu)
  int *csc_indptr = nullptr, *csc_indices = nullptr, *csc_map = nullptr;
  double* csc_data = nullptr;
  int csc_G = 0, csc_n = 0;
  // resident dense counts (sparse.cu): G rows of n cells, double (dtype 0) or int32 (dtype 1)
  void* dn_x = nullptr;
  int dn_dtype = 0, dn_G = 0, dn_n = 0;
};

inline int fail(saver_cuda_ctx* h, int code, const char* fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof h->err, fmt, ap);
    va_end(ap);
  }
  return code;
}

#define CU(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(h, e_ == cudaErrorMemoryAllocation ? SAVER_ERR_NOMEM : SAVER_ERR_CUDA, \
                  "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
  } while (0)

// Stream-ordered device buffer (cudaMallocAsync pool keeps freed blocks cached).
template <typename T>
struct DevBuf {
  T* p = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t alloc(size_t count, cudaStream_t stream) {
    s = stream;
    if (count == 0) count = 1;
    return cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), stream);
  }
  ~DevBuf() {
    if (p) cudaFreeAsync(p, s);
  }
};

// Input array that is either already on the device or must be uploaded.
template <typename T>
struct InArr {
  DevBuf<T> buf;
  const T* d = nullptr;
  cudaError_t bind(const T* src, size_t count, bool on_device, cudaStream_t s) {
    if (!src) { d = nullptr; return cudaSuccess; }
    if (on_device) { d = src; return cudaSuccess; }
    cudaError_t e = buf.alloc(count, s);
    if (e != cudaSuccess) return e;
    d = buf.p;
    return cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
  }
};

// Output array: device-resident caller buffer, or a temp that is copied back.
template <typename T>
struct OutArr {
  DevBuf<T> buf;
  T* d = nullptr;
  T* host = nullptr;
  size_t count = 0;
  cudaError_t bind(T* dst, size_t n, bool on_device, cudaStream_t s) {
    count = n;
    if (!dst) { d = nullptr; return cudaSuccess; }
    if (on_device) { d = dst; return cudaSuccess; }
    host = dst;
    cudaError_t e = buf.alloc(n, s);
    d = buf.p;
    return e;
  }
  cudaError_t fetch(cudaStream_t s) {
    if (!host || !d) return cudaSuccess;
    return cudaMemcpyAsync(host, d, count * sizeof(T), cudaMemcpyDeviceToHost, s);
  }
};
