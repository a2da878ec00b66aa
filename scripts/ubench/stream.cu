// stream.cu -- how fast can ONE CTA per SM stream design columns out of HBM?
//
// The Gram round kernel's streaming phases (strong-set check, entering Gram rows, Gram block) read
// design columns (n = 10,000 cells) picked per problem from a 0.7-1.5 GB matrix, one 512-thread CTA
// per SM with the working residual in shared memory.  This microbenchmark measures the dot
// g_c = sum_i wr_i x_ic for 64 random columns per "check" with the variants below, all 148 CTAs at
// once, and reports GB/s (the HBM peak is 6,552 GB/s = 44 GB/s per SM).
//
//   v0  warp per 4 columns, FP32 data -> FP64 math, predicated batches (wdot_a4f of lasso.cu)
//   v1  warp per 4 columns, FP32 data and FP32 math, unpredicated main loop
//   v2  warp per 4 columns, FP64 data (wdot_a4 of lasso.cu)
//   v3  block-cooperative: cp.async.bulk (TMA) slabs of 64 columns x R rows into a shared-memory
//       ring behind mbarriers, every warp owns 4 columns of the slab
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o stream stream.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int T = 512, NW = T / 32, NCOL = 64;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sumf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ unsigned lcg(unsigned& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// ---- v0 ----
__device__ __noinline__ void dot4_f32_f64math(const float* z0, const float* z1, const float* z2, const float* z3,
                                               const double* a, int ldx, double* out) {
  const int lane = threadIdx.x & 31;
  const float* zp[4] = {z0, z1, z2, z3};
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  constexpr int U = 3;
  for (int i = 4 * lane; i < ldx; i += 128 * U) {
    float4 zz[4][U];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        zz[c][u] = i + 128 * u < ldx ? __ldg((const float4*)(zp[c] + i + 128 * u)) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = i + 128 * u < ldx;
      const double2 ra = ok ? *(const double2*)(a + i + 128 * u) : make_double2(0, 0);
      const double2 rb = ok ? *(const double2*)(a + i + 128 * u + 2) : make_double2(0, 0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s0[c] += ra.x * (double)zz[c][u].x;
        s1[c] += ra.y * (double)zz[c][u].y;
        s0[c] += rb.x * (double)zz[c][u].z;
        s1[c] += rb.y * (double)zz[c][u].w;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = warp_sum(s0[c] + s1[c]);
}

// ---- v1: FP32 math, af = float copy of a in shared memory ----
template <int U>
__device__ __forceinline__ void dot4_f32(const float* z0, const float* z1, const float* z2, const float* z3,
                                         const float* af, int ldx, float* out) {
  const int lane = threadIdx.x & 31;
  const float* zp[4] = {z0, z1, z2, z3};
  float s[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) s[c][k] = 0.f;
  int i = 4 * lane;
  for (; i + 128 * (U - 1) < ldx; i += 128 * U) {
    float4 zz[4][U];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c) zz[c][u] = __ldg((const float4*)(zp[c] + i + 128 * u));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4 r = *(const float4*)(af + i + 128 * u);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s[c][0] = fmaf(r.x, zz[c][u].x, s[c][0]);
        s[c][1] = fmaf(r.y, zz[c][u].y, s[c][1]);
        s[c][2] = fmaf(r.z, zz[c][u].z, s[c][2]);
        s[c][3] = fmaf(r.w, zz[c][u].w, s[c][3]);
      }
    }
  }
  for (; i < ldx; i += 128) {
    const float4 r = *(const float4*)(af + i);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 z = __ldg((const float4*)(zp[c] + i));
      s[c][0] = fmaf(r.x, z.x, s[c][0]);
      s[c][1] = fmaf(r.y, z.y, s[c][1]);
      s[c][2] = fmaf(r.z, z.z, s[c][2]);
      s[c][3] = fmaf(r.w, z.w, s[c][3]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = warp_sumf((s[c][0] + s[c][1]) + (s[c][2] + s[c][3]));
}

// ---- v2: FP64 data ----
__device__ __noinline__ void dot4_f64(const double* z0, const double* z1, const double* z2, const double* z3,
                                      const double* a, int ldx, double* out) {
  const int lane = threadIdx.x & 31;
  const double* zp[4] = {z0, z1, z2, z3};
  double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
  constexpr int U = 4;
  for (int i = 2 * lane; i < ldx; i += 64 * U) {
    double2 zz[4][U];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        zz[c][u] = i + 64 * u < ldx ? __ldg((const double2*)(zp[c] + i + 64 * u)) : make_double2(0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const double2 r = i + 64 * u < ldx ? *(const double2*)(a + i + 64 * u) : make_double2(0, 0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        s0[c] += r.x * zz[c][u].x;
        s1[c] += r.y * zz[c][u].y;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c] = warp_sum(s0[c] + s1[c]);
}

// ---- mbarrier / bulk copy ----
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
          (unsigned)__cvta_generic_to_shared(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem)),
               "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b))
               : "memory");
}

template <int V, int ROWS, int STAGES>
__global__ void __launch_bounds__(T, 1)
stream_kernel(const float* __restrict__ Xf, const double* __restrict__ Xd, int ldx, int p, int checks,
              double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char sm[];
  double* a = (double*)sm;                      // ldx doubles
  float* af = (float*)(a + ldx);                // ldx floats
  int* cols = (int*)(af + ldx);                 // NCOL
  uint64_t* bar = (uint64_t*)(cols + NCOL);     // STAGES full barriers
  float* ring = (float*)(((uintptr_t)(bar + 8) + 127) / 128 * 128);  // STAGES x NCOL x ROWS floats
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < ldx; i += T) {
    a[i] = 1.0 / (1 + (i % 7));
    af[i] = (float)a[i];
  }
  if (V == 3 && tid == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(bar + s, 1);
  __syncthreads();
  unsigned seed = 12345u + blockIdx.x * 977u;
  double total = 0.0;
  unsigned it = 0;  // slab counter (ring position / parity) across checks
  for (int ck = 0; ck < checks; ++ck) {
    if (tid < NCOL) { unsigned s2 = seed + tid * 7919u + ck * 104729u; cols[tid] = lcg(s2) % p; }
    __syncthreads();
    if (V == 0) {
      const float* z[4];
      for (int c = 0; c < 4; ++c) z[c] = Xf + (size_t)cols[4 * warp + c] * ldx;
      double o[4];
      dot4_f32_f64math(z[0], z[1], z[2], z[3], a, ldx, o);
      total += o[0] + o[1] + o[2] + o[3];
    } else if (V == 1) {
      const float* z[4];
      for (int c = 0; c < 4; ++c) z[c] = Xf + (size_t)cols[4 * warp + c] * ldx;
      float o[4];
      dot4_f32<3>(z[0], z[1], z[2], z[3], af, ldx, o);
      total += (double)o[0] + o[1] + o[2] + o[3];
    } else if (V == 2) {
      const double* z[4];
      for (int c = 0; c < 4; ++c) z[c] = Xd + (size_t)cols[4 * warp + c] * ldx;
      double o[4];
      dot4_f64(z[0], z[1], z[2], z[3], a, ldx, o);
      total += o[0] + o[1] + o[2] + o[3];
    } else {
      // slabs of ROWS rows of all 64 columns; producer = lanes of warp 0 (2 columns each)
      const int nslab = (ldx + ROWS - 1) / ROWS;
      auto issue = [&](int sl, unsigned itn) {
        const int st = itn % STAGES;
        const int r0 = sl * ROWS, nr = ldx - r0 < ROWS ? ldx - r0 : ROWS;
        if (tid == 0) mbar_expect(bar + st, (unsigned)(NCOL * nr * 4));
        __syncwarp();
        for (int c = lane; c < NCOL; c += 32)
          bulk_g2s(ring + ((size_t)st * NCOL + c) * ROWS, Xf + (size_t)cols[c] * ldx + r0, (unsigned)(nr * 4), bar + st);
      };
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (warp == 0)
        for (int s = 0; s < STAGES - 1 && s < nslab; ++s) issue(s, it + s);
      for (int sl = 0; sl < nslab; ++sl, ++it) {
        // the stage refilled now was consumed in the previous iteration (barrier below)
        if (warp == 0 && sl + STAGES - 1 < nslab) issue(sl + STAGES - 1, it + STAGES - 1);
        const int st = it % STAGES;
        mbar_wait(bar + st, (it / STAGES) & 1);
        const int r0 = sl * ROWS, nr = ldx - r0 < ROWS ? ldx - r0 : ROWS;
        const float* rb = ring + ((size_t)st * NCOL + 4 * warp) * ROWS;
        for (int i = 4 * lane; i < nr; i += 128) {
          const float4 r = *(const float4*)(af + r0 + i);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 z = *(const float4*)(rb + c * ROWS + i);
            acc[c] = fmaf(r.x, z.x, acc[c]);
            acc[c] = fmaf(r.y, z.y, acc[c]);
            acc[c] = fmaf(r.z, z.z, acc[c]);
            acc[c] = fmaf(r.w, z.w, acc[c]);
          }
        }
        __syncthreads();  // everybody done with stage st before it is refilled
      }
      total += (double)warp_sumf(acc[0]) + warp_sumf(acc[1]) + warp_sumf(acc[2]) + warp_sumf(acc[3]);
    }
    __syncthreads();
  }
  if (lane == 0) atomicAdd(out + blockIdx.x, total);
}

template <int V, int ROWS, int STAGES>
void run(const char* name, const float* Xf, const double* Xd, int ldx, int p, int checks, double* out, double bytes_per_el) {
  size_t smem = (size_t)ldx * 12 + NCOL * 4 + 64 + 256 + (V == 3 ? (size_t)STAGES * NCOL * ROWS * 4 : 0);
  if (V != 3 && smem < 200 * 1024) smem = 200 * 1024;  // one CTA per SM like the round kernel
  CK(cudaFuncSetAttribute(stream_kernel<V, ROWS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaMemset(out, 0, 148 * 8));
    CK(cudaEventRecord(e0));
    stream_kernel<V, ROWS, STAGES><<<148, T, smem>>>(Xf, Xd, ldx, p, checks, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  double h[2];
  CK(cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost));
  const double bytes = 148.0 * checks * NCOL * (double)ldx * bytes_per_el;
  printf("%-44s smem %6zu KB  %8.2f ms  %7.1f GB/s  (%.1f GB/s per SM)  chk %.6g\n", name, smem / 1024, ms,
         bytes / ms / 1e6, bytes / ms / 1e6 / 148, h[0]);
}

__global__ void fill(float* Xf, double* Xd, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = (float)((i * 2654435761u) % 2001) / 1000.f - 1.f;
    Xf[i] = v;
    if (Xd) Xd[i] = v;
  }
}

int main(int argc, char** argv) {
  const int ldx = 10000, p = 18432, checks = argc > 1 ? atoi(argv[1]) : 40;
  float* Xf;
  double *Xd, *out;
  const size_t n = (size_t)ldx * p;
  CK(cudaMalloc(&Xf, n * 4));
  CK(cudaMalloc(&Xd, n * 8));
  CK(cudaMalloc(&out, 148 * 8));
  fill<<<148 * 8, 256>>>(Xf, Xd, n);
  CK(cudaDeviceSynchronize());
  run<0, 0, 1>("v0 warp x4 cols, f32 data, f64 math", Xf, Xd, ldx, p, checks, out, 4);
  run<1, 0, 1>("v1 warp x4 cols, f32 data, f32 math", Xf, Xd, ldx, p, checks, out, 4);
  run<2, 0, 1>("v2 warp x4 cols, f64 data", Xf, Xd, ldx, p, checks, out, 8);
  run<3, 64, 2>("v3 TMA ring 64 cols x 64 rows x 2 stages", Xf, Xd, ldx, p, checks, out, 4);
  run<3, 64, 3>("v3 TMA ring 64 cols x 64 rows x 3 stages", Xf, Xd, ldx, p, checks, out, 4);
  run<3, 128, 2>("v3 TMA ring 64 cols x 128 rows x 2 stages", Xf, Xd, ldx, p, checks, out, 4);
  run<3, 128, 3>("v3 TMA ring 64 cols x 128 rows x 3 stages", Xf, Xd, ldx, p, checks, out, 4);
  run<3, 32, 3>("v3 TMA ring 64 cols x 32 rows x 3 stages", Xf, Xd, ldx, p, checks, out, 4);
  return 0;
}
