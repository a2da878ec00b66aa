// gemm64.cu -- FP64 "TN" GEMM on the sm_100a DMMA path (mma.sync.m8n8k4.f64):
//     C[M x N] = A^T * B,   A: K x M column-major (lda),  B: K x N column-major (ldb).
//
// This is the shared-design contraction of the lasso path: A = the standardised design
// Xs (cells x genes, K = cells), B = a panel of per-problem working residuals (or of
// centred responses / fold weights), C = gradients X^T r for every (variable, problem).
// It replaces the p separate dot_product(wr, x(:,k)) loops that glmnet's fishnet1 runs per
// gene for its KKT check, strong rule and lambda_max (called from R/expr_predict.R:41,65),
// and stats::cor in R/calc_maxcor.R:27.
//
// tcgen05 has no FP64 kind, so the exact contraction uses DMMA; operands are staged through
// a 4-deep cp.async ring in shared memory.  A stage keeps every operand row's 16 k-values
// contiguous (eight 16-byte chunks, chunk index XORed with 4 on odd rows).  An m8n8k4 MMA may sum
// any four k's as long as A and B agree, so lane (g, t) takes the k PAIR {2t, 2t+1} of its row
// with ONE 16-byte LDS and feeds two MMAs (even k's, odd k's): 8 conflict-free LDS.128 per 32
// DMMA, half the shared-memory instructions of one LDS.64 per fragment.
//
// Caller guarantees (the library allocates its panels that way): lda, ldb multiples of 16
// with zero-filled padding rows up to Kp (multiple of 16); A has >= round_up(M,128) columns
// and B has >= round_up(N,64) columns allocated (contents of the padding columns are
// irrelevant, they only feed outputs that are not stored).
#include "common.cuh"
#include "saver_kernels.h"

namespace saver {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 4, GTHREADS = 256;
constexpr int A_STAGE = BM * BK;  // doubles
constexpr int B_STAGE = BN * BK;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(GTHREADS, 2)
gemm_tn_f64_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                   double* __restrict__ C, int ldc, int M, int N, int Kp) {
  extern __shared__ __align__(16) double sm[];
  double* As = sm;                     // [STAGES][BM][BK], chunks swizzled
  double* Bs = sm + STAGES * A_STAGE;  // [STAGES][BN][BK]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;  // 4 x 2 warps, 32 x 32 each
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const double* Ag = A + (size_t)m0 * lda;
  const double* Bg = B + (size_t)n0 * ldb;

  // chunk q (k = 2q, 2q+1) of operand row r lives at r * 16 + 2 * (q ^ ((r & 1) << 2))
  auto load_stage = [&](int stage, int k0) {
    double* as = As + stage * A_STAGE;
    double* bs = Bs + stage * B_STAGE;
#pragma unroll
    for (int i = 0; i < (BM * 8) / GTHREADS; ++i) {
      const int id = tid + GTHREADS * i, m = id >> 3, q = id & 7;
      cp_async16(as + m * BK + 2 * (q ^ ((m & 1) << 2)), Ag + (size_t)m * lda + k0 + 2 * q);
    }
#pragma unroll
    for (int i = 0; i < (BN * 8) / GTHREADS; ++i) {
      const int id = tid + GTHREADS * i, c = id >> 3, q = id & 7;
      cp_async16(bs + c * BK + 2 * (q ^ ((c & 1) << 2)), Bg + (size_t)c * ldb + k0 + 2 * q);
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int g8 = lane >> 2, t4 = lane & 3;
  const int swz = (g8 & 1) << 2;  // operand rows wm * 32 + i * 8 + g8 have the parity of g8
  const int KT = Kp / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_stage(s, s * BK);
    cp_commit();
  }
  for (int kt = 0; kt < KT; ++kt) {
    cp_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    if (nxt < KT) load_stage(nxt % STAGES, nxt * BK);
    cp_commit();
    const double* as = As + (kt % STAGES) * A_STAGE + (wm * 32 + g8) * BK;
    const double* bs = Bs + (kt % STAGES) * B_STAGE + (wn * 32 + g8) * BK;
#pragma unroll
    for (int kg = 0; kg < BK / 8; ++kg) {
      const int off = 2 * ((kg * 4 + t4) ^ swz);
      double2 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const double2*>(as + i * 8 * BK + off);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const double2*>(bs + j * 8 * BK + off);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
    }
  }
  cp_wait<0>();
  // epilogue: thread holds C[row = lane/4][col = 2*(lane%4) + {0,1}] of each 8x8 tile
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + wm * 32 + i * 8 + (lane >> 2);
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
      if (col < N) C[(size_t)col * ldc + row] = acc[i][j][0];
      if (col + 1 < N) C[(size_t)(col + 1) * ldc + row] = acc[i][j][1];
    }
  }
}

}  // namespace

cudaError_t launch_gemm_tn_f64(const double* A, int lda, const double* B, int ldb, double* C,
                               int ldc, int M, int N, int Kp, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  const int smem = STAGES * (A_STAGE + B_STAGE) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(gemm_tn_f64_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
  gemm_tn_f64_kernel<<<grid, GTHREADS, smem, stream>>>(A, lda, B, ldb, C, ldc, M, N, Kp);
  return cudaGetLastError();
}

}  // namespace saver
