// lfd_common.cuh -- shared declarations of the L_fd (cross-modal decorrelation loss) kernels.
//
// Replaces losses/fddm_losses.py:18-58 of the reference.  Data flow (z_a, z_b: [B,T,D]; rows = B*T):
//
//   lfd_stats      per (t,d) fp64 moments  sum_b x, sum_b x^2                       (one read of z_a, z_b)
//        | all-reduce SUM over ranks when the batch is sharded
//   lfd_xcov       scale/shift tables (rstd, -mean*rstd), then the tcgen05 contraction
//                  cov[j,k] = sum_rows za~[row,j] * zb~[row,k]   (split-K partials -> fixed-order sum)
//        | all-reduce SUM
//   lfd_loss       C = cov/N; loss = sum_j (1-C_jj)^2 + lambda sum_{j!=k} C_jk^2;  G = dloss/dC
//   lfd_backward   phase 0: dza~ = zb~ G^T/N, dzb~ = za~ G/N (two tcgen05 contractions), batch sums of
//                           dz~ and dz~*z~ per (t,d)        | all-reduce SUM
//                  phase 1: dx = (dz~ - mean_b dz~ - z~ mean_b(dz~ z~)) * rstd * upstream
#pragma once

#include "common.cuh"

namespace fddm {

// ---- workspace layout (bytes), shared by lfd_kernels.cu and lfd_umma.cu -------------------------
struct LfdWorkspace {
  static constexpr size_t kCounters = 256;            // self-resetting unsigned counters
  static constexpr size_t kMaxPartials = 1024;        // loss partial sums (double)
  static constexpr int kMaxSplits = 148;
  size_t off_partials, off_diag, off_tables, off_splitk, off_dza, off_dzb, total;
  __host__ __device__ LfdWorkspace(int64_t B, int64_t T, int64_t D) {
    const size_t td = static_cast<size_t>(T) * D, rows = static_cast<size_t>(B) * T;
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    off_partials = kCounters;
    off_diag = al(off_partials + kMaxPartials * sizeof(double));          // exact fp64 diagonal of cov
    off_tables = al(off_diag + static_cast<size_t>(D) * sizeof(double));
    off_splitk = al(off_tables + 4 * td * sizeof(float));               // a_scale a_shift b_scale b_shift
    off_dza = al(off_splitk + static_cast<size_t>(kMaxSplits) * D * D * sizeof(float));
    off_dzb = al(off_dza + rows * D * sizeof(float));
    total = al(off_dzb + rows * D * sizeof(float));
  }
};

// ---- the generic tcgen05 contraction (lfd_umma.cu) ----------------------------------------------
// An operand is a row-major global matrix X[r][c] (c contiguous, `ld` elements per row) of which the
// MMA sees element (mn, k) = mn_is_col ? X[k][mn] : X[mn][k], optionally standardised on the way
// into shared memory:  x~ = x * scale[(r % T) * stat_ld + c] + shift[...]   (scale == nullptr: raw).
struct UmmaOperand {
  const void* ptr;
  int dtype;             // fddm_dtype_t
  int64_t ld;
  int64_t nrows, ncols;  // extent of X
  int mn_is_col;
  const float* scale;
  const float* shift;
  int T;
  int64_t stat_ld;
};

// out[s][m][n] = alpha * sum_{k in split s} A(m,k) * B(n,k)      (fp32, m < M, n < N)
// terms: 1 = operands rounded to bf16; 2 = bf16 hi + bf16 residual (hi*hi + hi*lo + lo*hi, ~2^-16)
int umma_gemm(const UmmaOperand& A, const UmmaOperand& B, int64_t M, int64_t N, int64_t K, int splits, int terms,
              float alpha, float* out, int64_t out_ld, int64_t out_split_stride, cudaStream_t stream);

}  // namespace fddm
