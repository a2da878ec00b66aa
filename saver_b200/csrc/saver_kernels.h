// Internal launch interfaces between the C-ABI layer (saver_cuda.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "lasso.cuh"

namespace saver {

constexpr int kPostInfo = 8;  // per-gene info record of calc_post (doubles)

struct CalcPostArgs {
  const double* Y;    // n x B raw counts (device)
  const double* MU;   // n x B predictions (device)
  const double* sf;   // n size factors (device)
  int n, B;
  double scale_sf;
  double* EST;        // n x B (ceiling-quantised), required
  double* SE;         // n x B or null (estimates.only)
  double* EST_RAW;    // optional pre-ceiling outputs (parity tests)
  double* SE_RAW;
  double* INFO;       // kPostInfo x B: method, param, nll a/b/k, fallback mask, #evals, error mask
  double* GRID;       // optional 600 x B: per model 100 grid points + 100 log-liks
  int* nz_ws;         // n x B ints, only used when n is too large to stage (STAGE 0)
};

struct LogLikArgs {
  const double *y, *mu, *sf;  // device, length n
  int n, model, mode;         // mode 0: loglik(param); mode 1: calc.a/b/k
  double param;
  double* out;                // 4 doubles (device)
  double* grid;               // optional 200 doubles (device)
  int* nz_ws;                 // n ints (device)
};

// The shared design: x.est of R/saver.R:156-157, standardised once over all cells.
struct DesignState {
  int n = 0, p = 0;
  int ldx = 0;              // leading dimension (rows) of Xs: n rounded up to 16, zero padded
  int p_pad = 0;            // allocated columns: p rounded up to 128
  double* Xs = nullptr;     // ldx x p_pad column-major, (x - mean) / sd_n (1/n variance); constant columns -> 0
  double* Xs2 = nullptr;    // Xs squared elementwise (per-fit second moments come from one GEMM)
  float* Xf = nullptr;      // Xs rounded to FP32: feeds the lasso's curvature model only (Gram blocks)
  double* xmean = nullptr;  // p
  double* xsd = nullptr;    // p (0 marks a constant column)
  cudaError_t build(const double* X_dev, int n, int p, cudaStream_t s);
  void release() {
    cudaFree(Xs); cudaFree(Xs2); cudaFree(Xf); cudaFree(xmean); cudaFree(xsd);
    Xs = Xs2 = xmean = xsd = nullptr;
    Xf = nullptr;
    n = p = ldx = p_pad = 0;
  }
};

// C[M x N] (ldc) = A^T B; A: Kp x M (lda), B: Kp x N (ldb), column-major, see gemm64.cu
cudaError_t launch_gemm_tn_f64(const double* A, int lda, const double* B, int ldb, double* C,
                               int ldc, int M, int N, int Kp, cudaStream_t stream);
cudaError_t launch_center_scale(const double* Y, int n, int ldr, int B, double* R, cudaStream_t s);
cudaError_t launch_colmax(const double* C, int ldc, int p, int B, const int* self, double* out,
                          cudaStream_t s);

// Batched Poisson-lasso paths of one block of genes (lasso.cu).  Device pointers throughout.
int lasso_max_block(int NF, const DesignState& D, const LassoTuning& tune);
cudaError_t lasso_run(const DesignState& D, LassoWs& ws, const LassoParams& prm, const double* Yn,
                      const int* fold, const int* self, const double* lam_user,
                      const int* nlam_user, const double* lmin_user, double* MU, double* out4,
                      int* status, int* stats, double* cvout, cudaStream_t s, long long* launches,
                      long long* gemm_cols);

cudaError_t lasso_debug_path(const LassoWs& ws, int g, double* a0, double* dev, int* nin, int* nlp,
                             double* lam);

size_t calc_post_smem_bytes(int n, int* stage);
cudaError_t launch_calc_post(const CalcPostArgs& A, cudaStream_t stream);
cudaError_t launch_loglik(const LogLikArgs& A, cudaStream_t stream);

}  // namespace saver
