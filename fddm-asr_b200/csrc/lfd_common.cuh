// lfd_common.cuh -- shared declarations of the L_fd (cross-modal decorrelation loss) kernels.
//
// Replaces losses/fddm_losses.py:18-58 of the reference.  Data flow (z_a, z_b: [B,T,D]; rows = B*T):
//
//   lfd_stats      per (t,d) fp64 moments  sum_b x, sum_b x^2                       (one read of z_a, z_b)
//        | all-reduce SUM over ranks when the batch is sharded
//   lfd_xcov       scale/shift tables (rstd, -mean*rstd); ONE pass over z_a, z_b that standardises, splits
//                  into bf16 hi + residual and writes the tensor-core operand planes in "packed" form, and
//                  accumulates the exact fp64 diagonal; then the tcgen05 contraction
//                  cov[j,k] = sum_rows za~[row,j] * zb~[row,k]   (split-K partials -> fixed-order sum)
//        | all-reduce SUM
//   lfd_loss       C = cov/N; loss = sum_j (1-C_jj)^2 + lambda sum_{j!=k} C_jk^2;  G = dloss/dC
//   lfd_backward   phase 0: pack G and G^T; dza~ = zb~ G^T/N, dzb~ = za~ G/N (two tcgen05 contractions that
//                           re-use the packed planes), batch sums of dz~ and dz~*z~ per (t,d)
//                           | all-reduce SUM
//                  phase 1: dx = (dz~ - mean_b dz~ - z~ mean_b(dz~ z~)) * rstd * upstream
//
// Packed operand format P(X) of a row-major matrix X[R][C]: bf16 planes (hi, lo) laid out as
//     P[c/8][r][c%8]        (R and C zero-padded to multiples of kPackPad)
// i.e. for each group of 8 consecutive columns ("chunk column") all rows are contiguous, 16 bytes per
// row.  A run of rows of one chunk column is exactly a column of UMMA "core matrices" (8 rows x 16
// bytes = 128 contiguous bytes) of the canonical no-swizzle shared-memory layout, for BOTH operand
// orientations:
//   * MN index = column, K index = row (the forward z~^T z~): tile = BK rows x (TM/8) chunk columns,
//     MN-major descriptor, SBO = BK*16 (between chunk columns), LBO = 128 (between 8-row K groups)
//   * MN index = row, K index = column (the backward z~ G):   tile = TM rows x (BK/8) chunk columns,
//     K-major descriptor, SBO = 128 (between 8-row MN groups), LBO = TM*16 (between chunk columns)
// so the contraction kernel's producer is nothing but 1-D TMA bulk copies (cp.async.bulk) of
// contiguous runs into shared memory -- no tensor map, no swizzle, no register staging.
#pragma once

#include "common.cuh"

namespace fddm {

constexpr int kPackPad = 256;      // packed planes are zero-padded to multiples of this in R and C

__host__ __device__ inline int64_t pack_pad(int64_t x) { return (x + kPackPad - 1) / kPackPad * kPackPad; }

// Row order of the packed z~ planes.
//   B >= 32 ("tb-major"): packed row r = t * Bp + b with the batch padded to Bp = ceil32(B) (rows b >= B are
//     zero), so 32 consecutive packed rows -- the rows one epilogue warp of the backward contraction owns --
//     always belong to ONE position t: the batch sums of dz~ * z~ the batch-norm backward needs come out of the
//     contraction's epilogue as [T][Bp/32][D] partials instead of a separate pass over dz~ and z.
//   B < 32 (the reference's own configs use B = 2..16): natural order r = b * T + t, nothing is padded, and the
//     batch sums are taken by lfd_bn_reduce_kernel.
struct LfdRows {
  int B, T, Bp;
  bool tb_major;
  int64_t rows_packed;     // T * Bp (tb-major) or B * T
  __host__ __device__ LfdRows(int64_t B_, int64_t T_) {
    B = static_cast<int>(B_); T = static_cast<int>(T_);
    tb_major = B_ >= 32;
    Bp = tb_major ? static_cast<int>((B_ + 31) / 32 * 32) : B;
    rows_packed = tb_major ? T_ * Bp : B_ * T_;
  }
  __host__ __device__ int parts() const { return tb_major ? Bp / 32 : 1; }   // bn partials per (t, d)
};

struct PackedOperand {
  const __nv_bfloat16* hi;
  const __nv_bfloat16* lo;   // residual plane (may be null when terms == 1)
  int64_t R_pad, C_pad;      // padded extents of X
  int mn_is_col;             // 1: MMA (mn, k) = X[k][mn];  0: MMA (mn, k) = X[mn][k]
};

// out[s][m][n] = alpha * sum_{k in split s} A(m,k) * B(n,k)      (fp32, m < M, n < N)
// terms: 1 = hi planes only; 2 = hi*hi + hi*lo + lo*hi (~2^-16 relative)
int umma_gemm(const PackedOperand& A, const PackedOperand& B, int64_t M, int64_t N, int64_t K, int splits, int terms,
              float alpha, float* out, int64_t out_ld, int64_t out_split_stride, cudaStream_t stream);

// Backward contraction for tb-major planes (lfd_umma_bwd.cu), persistent CTAs, two TMEM accumulators:
//   dz[b][t][n] = alpha * sum_k Z(r, k) * G(n, k)                 r = t * Bp + b, fp32, natural [B][T][D] order
//   partial[r / 32][n] = sum over the 32 rows of dz * Zs(r, n)     Zs = the z~ planes of the tensor dz belongs to
// Z, Zs: packed z~ planes (hi + lo), G: packed planes of G or G^T (hi + lo); N = K = D.
int umma_bwd_gemm(const PackedOperand& Z, const PackedOperand& G, const PackedOperand& Zs, const LfdRows& rows, int64_t D,
                  float alpha, float* dz, float* partial, cudaStream_t stream);

// ---- workspace layout (bytes), shared by lfd_kernels.cu and lfd_umma.cu -------------------------
struct LfdWorkspace {
  static constexpr size_t kCounters = 256;            // self-resetting unsigned counters
  static constexpr size_t kMaxPartials = 1024;        // loss partial sums (double)
  static constexpr int kMaxSplits = 32;
  size_t off_partials, off_diag, off_tables, off_splitk, off_pack, off_gpack, off_dza, off_dzb, total;
  size_t plane_bytes, gplane_bytes;
  __host__ __device__ LfdWorkspace(int64_t B, int64_t T, int64_t D) {
    const size_t td = static_cast<size_t>(T) * D, rows = static_cast<size_t>(B) * T;
    auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
    const size_t Rp = static_cast<size_t>(pack_pad(LfdRows(B, T).rows_packed)), Dp = static_cast<size_t>(pack_pad(D));
    plane_bytes = Rp * Dp * 2;
    gplane_bytes = Dp * Dp * 2;
    off_partials = kCounters;
    off_diag = al(off_partials + kMaxPartials * sizeof(double));          // exact fp64 diagonal of cov
    off_tables = al(off_diag + static_cast<size_t>(D) * sizeof(double));
    off_splitk = al(off_tables + 4 * td * sizeof(float));                 // a_scale a_shift b_scale b_shift
    off_pack = al(off_splitk + static_cast<size_t>(kMaxSplits) * D * D * sizeof(float));
    off_gpack = al(off_pack + 4 * plane_bytes);                           // za~ hi, lo, zb~ hi, lo
    off_dza = al(off_gpack + 4 * gplane_bytes);                           // G hi, lo, G^T hi, lo
    off_dzb = al(off_dza + rows * D * sizeof(float));
    total = al(off_dzb + rows * D * sizeof(float));
  }
};

}  // namespace fddm
