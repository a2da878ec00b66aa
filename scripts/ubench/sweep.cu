// Stand-alone timing of the Gram-space warp sweeps of lasso.cu (same functions, synthetic block).
#include "../../saver_b200/csrc/lasso.cu"
#include <cstdio>
#include <vector>
using namespace saver;

template <int Q, bool SM>
__global__ void __launch_bounds__(256, 2) sweep_bench(const double* Hg, int m, int reps, long long* cyc, double* sink, int busy) {
  extern __shared__ __align__(16) double sm[];
  double* gq = sm; double* va = gq + 640; double* vinv = va + 640; double* aact = vinv + 640; double* swa = aact + 640;
  double* gsh = swa + 640; int* ord = reinterpret_cast<int*>(gsh + 8); double* Hp = reinterpret_cast<double*>(ord + 640);
  const int tid = threadIdx.x;
  for (int l = tid; l < m; l += 256) {
    va[l] = Hg[l + (size_t)l * 640]; vinv[l] = 1.0 / va[l]; aact[l] = 0.0; swa[l] = 0.01; gq[l] = 0.5 + 0.01 * l; ord[l] = (l * 37) % m;
  }
  if (SM) for (int idx = tid; idx < m * m; idx += 256) { const int r = idx % m, c = idx / m; if (r >= c) Hp[r * (r + 1) / 2 + c] = Hg[r + (size_t)c * 640]; }
  __syncthreads();
  if (blockIdx.x == 0 || !busy) {
    long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
      if (tid < 32) {
        if (SM) gram_sweep_sm<Q>(Hp, m, it & 1, ord, gq, va, vinv, aact, swa, 0.3, gsh, 0.0, 0.0);
        else gram_sweep_warp<4, 4>(Hg, 640, m, it & 1, ord, gq, va, vinv, aact, swa, 0.3, gsh, 0.0, 0.0);
      }
      __syncthreads();
      if (gsh[2] == 123.0) aact[0] = 1;
      __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = (t1 - t0);
  } else {  // neighbours hammering the FP64 pipe with DMMA
    double d0 = 0, d1 = 0;
    for (int i = 0; i < 200000; ++i) dmma(d0, d1, 1.0 + tid, 0.5);
    sink[blockIdx.x * 256 + tid] = d0 + d1;
  }
  if (tid == 0) sink[blockIdx.x] = gq[0] + aact[1];
}

int main() {
  const int m = 79, reps = 50;
  std::vector<double> H(640 * 640, 0.0);
  for (int r = 0; r < m; ++r) for (int c = 0; c < m; ++c) H[r + 640 * c] = (r == c) ? 1.0 : 0.3 / (1 + abs(r - c));
  double* dH; long long* dc; double* ds;
  cudaMalloc(&dH, H.size() * 8); cudaMalloc(&dc, 8 * 512); cudaMalloc(&ds, 8 * 512 * 256);
  cudaMemcpy(dH, H.data(), H.size() * 8, cudaMemcpyHostToDevice);
  const size_t smem = 5 * 640 * 8 + 64 + 640 * 4 + (size_t)m * (m + 1) / 2 * 8 + 64;
  cudaFuncSetAttribute(sweep_bench<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(sweep_bench<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long h;
  for (int busy = 0; busy < 2; ++busy) {
    const int grid = busy ? 296 : 1;
    for (int rep = 0; rep < 2; ++rep) {
      sweep_bench<3, true><<<grid, 256, smem, 0>>>(dH, m, reps, dc, ds, busy);
      cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
      printf("smem sweep   busy=%d : %.1f cycles/step (%s)\n", busy, (double)h / (reps * m), cudaGetErrorString(cudaGetLastError()));
      sweep_bench<3, false><<<grid, 256, smem, 0>>>(dH, m, reps, dc, ds, busy);
      cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
      printf("global sweep busy=%d : %.1f cycles/step (%s)\n", busy, (double)h / (reps * m), cudaGetErrorString(cudaGetLastError()));
    }
  }
  // large smem carve-out (little L1), as in the round kernel
  sweep_bench<3, false><<<1, 256, 100 * 1024, 0>>>(dH, m, reps, dc, ds, 0);
  cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
  printf("global sweep, 100 KB smem : %.1f cycles/step\n", (double)h / (reps * m));
  return 0;
}
