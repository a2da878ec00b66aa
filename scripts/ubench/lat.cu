// Latency microbenchmarks for the serial chain of a Gram-space coordinate step (B200, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, long long* cyc, double a, double b, int idx) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 1023);
  __syncthreads();
  double x = a + threadIdx.x, y = b;
  double d0 = 0, d1 = 0;
  int j = idx;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (MODE == 0) x = fma(x, y, a);
    if (MODE == 1) x = x + y;
    if (MODE == 2) x = x * y;
    if (MODE == 3) x = __shfl_sync(0xffffffffu, x, (idx + i) & 31);
    if (MODE == 4) { j = (int)sm[j]; }
    if (MODE == 5) { x = (fabs(x) - a > 0.0) ? x * y : x + y; }
    if (MODE == 6) dmma(d0, d1, x, y);
    if (MODE == 7) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = f; }
    if (MODE == 8) { float f = __double2float_rn(x); x = (double)f + y; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + j + d0 + d1;
}
template <int MODE> void run(const char* name, int threads, int blocks) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * threads * blocks); cudaMalloc(&cyc, 8 * blocks);
  k<MODE><<<blocks, threads>>>(out, cyc, 1.0000001, 0.9999999, 3);
  k<MODE><<<blocks, threads>>>(out, cyc, 1.0000001, 0.9999999, 3);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads %4d blocks %3d : %.1f cycles/op (per warp)\n", name, threads, blocks, (double)h / N);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int t : {32, 128, 256, 512}) {
    run<0>("DFMA dependent", t, 1);
    run<1>("DADD dependent", t, 1);
    run<2>("DMUL dependent", t, 1);
    run<6>("DMMA dependent", t, 1);
  }
  run<3>("SHFL dependent", 32, 1);
  run<4>("LDS pointer chase (f64->int)", 32, 1);
  run<5>("DSETP+select+DMUL/DADD", 32, 1);
  run<7>("f64->f32 FFMA f32->f64", 32, 1);
  run<8>("f64->f32->f64 + DADD", 32, 1);
  run<0>("DFMA dependent, 2 CTAs/SM", 256, 296);
  run<6>("DMMA dependent, 2 CTAs/SM", 256, 296);
  return 0;
}
