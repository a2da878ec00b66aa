// Shared device helpers for libsaver_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

namespace saver {

constexpr int kWarp = 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of K values per thread.  Every thread returns the same bits
// (the cross-warp combine is done by every thread in a fixed order), so scalar
// control flow that depends on the result stays uniform across the CTA.
// scratch: K * (blockDim.x / 32) doubles of shared memory.
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    v[k] = warp_sum(v[k]);
    if (lane == 0) scratch[k * nw + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double s = 0;
    for (int w = 0; w < nw; ++w) s += scratch[k * nw + w];
    v[k] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ double block_sum1(double v, double* scratch) {
  double a[1] = {v};
  block_sum<1>(a, scratch);
  return a[0];
}

__device__ __forceinline__ double block_max1(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = scratch[0];
  for (int w = 1; w < nw; ++w) s = fmax(s, scratch[w]);
  __syncthreads();
  return s;
}
__device__ __forceinline__ double block_min1(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_min(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = scratch[0];
  for (int w = 1; w < nw; ++w) s = fmin(s, scratch[w]);
  __syncthreads();
  return s;
}

}  // namespace saver
