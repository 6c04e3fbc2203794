// lfd_kernels.cu -- L_fd (reference losses/fddm_losses.py:18-58): batch statistics, loss / dloss/dC,
// batch-norm backward, and the C entry points.  The D x D contractions live in lfd_umma.cu.
//
// Memory-bound pieces, algorithmic traffic per element of z (s = bytes per input element):
//   stats        read s            (both tensors in one launch)
//   bn_reduce    read s + 4        (z and the fp32 dz~ written by the backward contraction)
//   bn_finalize  read s + 4, write s
#include <math.h>

#include <algorithm>

#include "lfd_common.cuh"

namespace fddm {
namespace {

// ------------------------------------------------------------------------------------------------
// per-(t,d) moments over the batch axis.  z: [B][TD]; thread = one 16-byte vector of the TD plane.
// grid = (ceil(TD/N/256), ntensors, bsplit).  fp64 accumulation: raw moments without cancellation
// trouble, and an exact-enough SUM all-reduce when the batch is sharded.
// ------------------------------------------------------------------------------------------------
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) lfd_stats_kernel(const T* __restrict__ za, const T* __restrict__ zb, int B,
                                                        int64_t TD, int bsplit, double* __restrict__ sums) {
  constexpr int N = VEC ? Vec16<T>::N : 1;
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v * N >= TD) return;
  const T* z = (blockIdx.y == 0) ? za : zb;
  const int bchunk = (B + bsplit - 1) / bsplit;
  const int b0 = blockIdx.z * bchunk, b1 = min(B, b0 + bchunk);
  double s[N], q[N];
#pragma unroll
  for (int e = 0; e < N; ++e) { s[e] = 0.0; q[e] = 0.0; }
  constexpr int U = 4;
  int b = b0;
  for (; b + U <= b1; b += U) {
    float x[U][N];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const T* p = z + static_cast<int64_t>(b + u) * TD + v * N;
      if (VEC) Vec16<T>::unpack(ldg_stream_v4(p), x[u]);
      else x[u][0] = Vec16<T>::load1(p);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const double d = static_cast<double>(x[u][e]);
        s[e] += d;
        q[e] = fma(d, d, q[e]);
      }
    }
  }
  for (; b < b1; ++b) {
    float x[N];
    const T* p = z + static_cast<int64_t>(b) * TD + v * N;
    if (VEC) Vec16<T>::unpack(ldg_stream_v4(p), x);
    else x[0] = Vec16<T>::load1(p);
#pragma unroll
    for (int e = 0; e < N; ++e) {
      const double d = static_cast<double>(x[e]);
      s[e] += d;
      q[e] = fma(d, d, q[e]);
    }
  }
  double* out_s = sums + static_cast<int64_t>(blockIdx.y) * 2 * TD + v * N;
  double* out_q = out_s + TD;
#pragma unroll
  for (int e = 0; e < N; ++e) {
    if (bsplit == 1) { out_s[e] = s[e]; out_q[e] = q[e]; }
    else { atomicAdd(out_s + e, s[e]); atomicAdd(out_q + e, q[e]); }
  }
}

// Same moments, one CTA per (position t, strip of CW column vectors, tensor): thread (tx, ty) owns column vector
// tx and the batch rows ty, ty+RG, ...; the RG row groups are combined through shared memory in a fixed order
// and written directly -- no atomics, no memset, and per thread U 16-byte loads in flight at a stride pattern DRAM
// likes.  CW * RG = 256 threads.  <64, 4, 4> for long batches; <32, 8, 8> for B <= 128, where a thread's whole
// share of the batch (<= 16 rows) is then at most two rounds of loads instead of four and there are twice as
// many CTAs to hide the round trips (measured at B=64: 29-33 us against 31-38 us, i.e. within the run-to-run
// spread; the fixed ~14 us of this pass is launch ramp and tail, not the load depth).
// Needs D % N == 0 and 16-byte aligned inputs.  grid = (ceil(D/(CW N)), T, 2).
template <typename T, int CW, int RG, int U>
__global__ void __launch_bounds__(256) lfd_stats_rows_kernel(const T* __restrict__ za, const T* __restrict__ zb, int B,
                                                             int Tn, int D, double* __restrict__ sums) {
  constexpr int N = Vec16<T>::N;
  static_assert(CW * RG == 256, "256 threads");
  __shared__ double s_part[RG - 1][2][CW][N];
  const int tx = threadIdx.x % CW, ty = threadIdx.x / CW;
  const int t = blockIdx.y;
  const int c = (blockIdx.x * CW + tx) * N;
  const T* z = (blockIdx.z == 0 ? za : zb) + static_cast<int64_t>(t) * D + c;
  const int64_t bstride = static_cast<int64_t>(Tn) * D;
  double s[N], q[N];
#pragma unroll
  for (int e = 0; e < N; ++e) { s[e] = 0.0; q[e] = 0.0; }
  if (c < D) {
    int b = ty;
    for (; b + RG * (U - 1) < B; b += RG * U) {
      float x[U][N];
#pragma unroll
      for (int u = 0; u < U; ++u) Vec16<T>::unpack(ldg_stream_v4(z + (b + RG * u) * bstride), x[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const double d = static_cast<double>(x[u][e]);
          s[e] += d;
          q[e] = fma(d, d, q[e]);
        }
      }
    }
    for (; b < B; b += RG) {
      float x[N];
      Vec16<T>::unpack(ldg_stream_v4(z + b * bstride), x);
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const double d = static_cast<double>(x[e]);
        s[e] += d;
        q[e] = fma(d, d, q[e]);
      }
    }
  }
  if (ty > 0) {
#pragma unroll
    for (int e = 0; e < N; ++e) { s_part[ty - 1][0][tx][e] = s[e]; s_part[ty - 1][1][tx][e] = q[e]; }
  }
  __syncthreads();
  if (ty == 0 && c < D) {
    const int64_t TD = static_cast<int64_t>(Tn) * D;
    double* out_s = sums + static_cast<int64_t>(blockIdx.z) * 2 * TD + static_cast<int64_t>(t) * D + c;
#pragma unroll
    for (int e = 0; e < N; ++e) {
      double a = s[e], b2 = q[e];
#pragma unroll
      for (int g = 0; g < RG - 1; ++g) { a += s_part[g][0][tx][e]; b2 += s_part[g][1][tx][e]; }   // fixed order
      out_s[e] = a;
      out_s[TD + e] = b2;
    }
  }
}

// sums -> (scale = rstd, shift = -mean*rstd) fp32 tables for both tensors: tables[4][TD]
__global__ void __launch_bounds__(256) lfd_tables_kernel(const double* __restrict__ sums, int64_t TD, double n_batch,
                                                         double eps, float* __restrict__ tables) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 2 * TD) return;
  const int which = static_cast<int>(i / TD);
  const int64_t j = i - which * TD;
  const double mean = sums[which * 2 * TD + j] / n_batch;
  double var = sums[which * 2 * TD + TD + j] / n_batch - mean * mean;      // biased (losses:24)
  var = var < 0.0 ? 0.0 : var;
  const double rstd = 1.0 / sqrt(var + eps);                                // losses:25
  tables[which * 2 * TD + j] = static_cast<float>(rstd);
  tables[which * 2 * TD + TD + j] = static_cast<float>(-mean * rstd);
}

// One pass over z_a and z_b: standardise (losses:23-26), split into bf16 hi + residual, write the
// packed tensor-core operand planes P[c/8][r][c%8] (lfd_common.cuh), zero the padding, and accumulate
// the diagonal of cov in fp64:  diag[j] += sum_rows za~[r,j] zb~[r,j].
// Why the diagonal is done here on the CUDA cores: the tensor-core accumulator adds with truncation,
// harmless for the zero-mean off-diagonal sums but a systematic ~1e-6 relative bias on the all-positive
// diagonal sums when z_a and z_b are correlated -- and the loss is dominated by sum_j (1-C_jj)^2.
// Layout change through shared memory: phase A reads a 32-row x 64-column tile row-contiguously
// (256-byte runs, the standardisation is applied here), phase B has warp -> chunk column, lane -> row,
// so every packed-plane store is a contiguous 512-byte run.  Rows are padded to 68 floats: both the
// 16-byte stores of phase A and the 2 x 16-byte loads of phase B are bank-conflict free.
// grid = (D_pad/64, R_pad/256); a CTA walks 8 row tiles and issues its 64 x 1 fp64 atomics once.
constexpr int kPackLd = 68;

template <typename T>
__device__ __forceinline__ void pack_phase_a(const T* __restrict__ z, const float* __restrict__ scale,
                                             const float* __restrict__ shift, int64_t r0, int c0, int64_t rows, int Tn,
                                             int D, float (*tile)[kPackLd]) {
  constexpr int N = Vec16<T>::N;                 // elements per 16-byte global vector
  constexpr int VPR = 64 / N;                    // vectors per tile row
  for (int idx = threadIdx.x; idx < 32 * VPR; idx += 256) {
    const int row = idx / VPR, cv = idx % VPR;
    const int64_t r = r0 + row;
    const int c = c0 + cv * N;
    float x[N];
    if (r < rows && c < D) {
      Vec16<T>::unpack(ldg_stream_v4(z + r * D + c), x);
      const int64_t so = (r % Tn) * D + c;
#pragma unroll
      for (int q = 0; q < N / 4; ++q) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + so) + q);
        const float4 h4 = __ldg(reinterpret_cast<const float4*>(shift + so) + q);
        x[4 * q + 0] = fmaf(x[4 * q + 0], s4.x, h4.x); x[4 * q + 1] = fmaf(x[4 * q + 1], s4.y, h4.y);
        x[4 * q + 2] = fmaf(x[4 * q + 2], s4.z, h4.z); x[4 * q + 3] = fmaf(x[4 * q + 3], s4.w, h4.w);
      }
    } else {
#pragma unroll
      for (int e = 0; e < N; ++e) x[e] = 0.0f;
    }
#pragma unroll
    for (int q = 0; q < N / 4; ++q)
      *reinterpret_cast<float4*>(&tile[row][cv * N + 4 * q]) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) lfd_pack_kernel(const T* __restrict__ za, const T* __restrict__ zb,
                                                       const float* __restrict__ tables, int64_t rows, int Tn, int D,
                                                       int64_t R_pad, __nv_bfloat16* __restrict__ a_hi,
                                                       __nv_bfloat16* __restrict__ a_lo, __nv_bfloat16* __restrict__ b_hi,
                                                       __nv_bfloat16* __restrict__ b_lo, double* __restrict__ diag) {
  __shared__ __align__(16) float sA[32][kPackLd];
  __shared__ __align__(16) float sB[32][kPackLd];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 64;
  const int cc = blockIdx.x * 8 + warp;                      // this warp's chunk column in phase B
  const int64_t TD = static_cast<int64_t>(Tn) * D;
  double acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0;
  for (int it = 0; it < 8; ++it) {
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 256 + it * 32;
    pack_phase_a<T>(za, tables, tables + TD, r0, c0, rows, Tn, D, sA);
    pack_phase_a<T>(zb, tables + 2 * TD, tables + 3 * TD, r0, c0, rows, Tn, D, sB);
    __syncthreads();
    float a[8], b[8];
    {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[lane][warp * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[lane][warp * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sB[lane][warp * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sB[lane][warp * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
    }
    __syncthreads();                                         // tiles are rewritten by the next iteration
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += static_cast<double>(a[e]) * static_cast<double>(b[e]);   // pads are 0
    const int64_t off = (static_cast<int64_t>(cc) * R_pad + r0 + lane) * 8;                        // elements
    const uint4 ah = Vec16<__nv_bfloat16>::pack(a), bh = Vec16<__nv_bfloat16>::pack(b);
    stg_stream_v4(a_hi + off, ah);
    stg_stream_v4(b_hi + off, bh);
    if (a_lo != nullptr) {
      float t[8];
      Vec16<__nv_bfloat16>::unpack(ah, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = a[e] - t[e];                     // exact in fp32
      stg_stream_v4(a_lo + off, Vec16<__nv_bfloat16>::pack(t));
      Vec16<__nv_bfloat16>::unpack(bh, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = b[e] - t[e];
      stg_stream_v4(b_lo + off, Vec16<__nv_bfloat16>::pack(t));
    }
  }
  if (cc * 8 < D) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      double v = acc[e];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) atomicAdd(diag + cc * 8 + e, v);
    }
  }
}

// tb-major variant (B >= 32, lfd_common.cuh): packed row r = t * Bp + b.  A CTA owns one position t and one
// 64-column strip and walks the batch 32 rows at a time, so every thread touches the SAME columns in every
// iteration: the standardisation scale/shift of (t, column) are loaded once per CTA into registers instead of
// once per element.  Rows b >= B (batch padding) and, for the last t, the rows up to R_pad are written as zeros.
// grid = (D_pad/64, T).
template <typename T>
__global__ void __launch_bounds__(256) lfd_pack_tb_kernel(const T* __restrict__ za, const T* __restrict__ zb,
                                                          const float* __restrict__ tables, int Bn, int Bp, int Tn, int D,
                                                          int64_t R_pad, __nv_bfloat16* __restrict__ a_hi,
                                                          __nv_bfloat16* __restrict__ a_lo, __nv_bfloat16* __restrict__ b_hi,
                                                          __nv_bfloat16* __restrict__ b_lo, double* __restrict__ diag) {
  __shared__ __align__(16) float sA[32][kPackLd];
  __shared__ __align__(16) float sB[32][kPackLd];
  constexpr int N = Vec16<T>::N;                 // elements per 16-byte global vector
  constexpr int VPR = 64 / N;                    // vectors per tile row
  constexpr int RPI = 256 / VPR;                 // tile rows covered by one sweep of the 256 threads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.y;
  const int c0 = blockIdx.x * 64;
  const int cc = blockIdx.x * 8 + warp;          // this warp's chunk column in phase B
  const int64_t TD = static_cast<int64_t>(Tn) * D;
  // phase-A role of this thread: column vector cv of the strip (fixed), tile rows prow, prow + RPI, ...
  const int cv = threadIdx.x % VPR, prow = threadIdx.x / VPR;
  const int c = c0 + cv * N;
  const bool col_ok = c < D;
  float sca[N], sha[N], scb[N], shb[N];
#pragma unroll
  for (int e = 0; e < N; ++e) {
    const int64_t so = static_cast<int64_t>(t) * D + c + e;
    sca[e] = col_ok ? __ldg(tables + so) : 0.0f;
    sha[e] = col_ok ? __ldg(tables + TD + so) : 0.0f;
    scb[e] = col_ok ? __ldg(tables + 2 * TD + so) : 0.0f;
    shb[e] = col_ok ? __ldg(tables + 3 * TD + so) : 0.0f;
  }
  double acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0;
  const int64_t rbase = static_cast<int64_t>(t) * Bp;
  // the last position also zero-fills the plane rows between T*Bp and R_pad
  const int nrows = (t == Tn - 1) ? static_cast<int>(R_pad - rbase) : Bp;
  for (int b0 = 0; b0 < nrows; b0 += 32) {
#pragma unroll
    for (int rr = prow; rr < 32; rr += RPI) {
      const int b = b0 + rr;
      float xa[N], xb[N];
      if (b < Bn && col_ok) {
        const int64_t off = (static_cast<int64_t>(b) * Tn + t) * D + c;
        Vec16<T>::unpack(ldg_stream_v4(za + off), xa);
        Vec16<T>::unpack(ldg_stream_v4(zb + off), xb);
#pragma unroll
        for (int e = 0; e < N; ++e) { xa[e] = fmaf(xa[e], sca[e], sha[e]); xb[e] = fmaf(xb[e], scb[e], shb[e]); }
      } else {
#pragma unroll
        for (int e = 0; e < N; ++e) { xa[e] = 0.0f; xb[e] = 0.0f; }
      }
#pragma unroll
      for (int q = 0; q < N / 4; ++q) {
        *reinterpret_cast<float4*>(&sA[rr][cv * N + 4 * q]) = make_float4(xa[4 * q], xa[4 * q + 1], xa[4 * q + 2], xa[4 * q + 3]);
        *reinterpret_cast<float4*>(&sB[rr][cv * N + 4 * q]) = make_float4(xb[4 * q], xb[4 * q + 1], xb[4 * q + 2], xb[4 * q + 3]);
      }
    }
    __syncthreads();
    float a[8], b[8];
    {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[lane][warp * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[lane][warp * 8 + 4]);
      const float4 b0v = *reinterpret_cast<const float4*>(&sB[lane][warp * 8]);
      const float4 b1v = *reinterpret_cast<const float4*>(&sB[lane][warp * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0v.x; b[1] = b0v.y; b[2] = b0v.z; b[3] = b0v.w; b[4] = b1v.x; b[5] = b1v.y; b[6] = b1v.z; b[7] = b1v.w;
    }
    __syncthreads();                                         // tiles are rewritten by the next iteration
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += static_cast<double>(a[e]) * static_cast<double>(b[e]);   // pads are 0
    const int64_t off = (static_cast<int64_t>(cc) * R_pad + rbase + b0 + lane) * 8;                // elements
    const uint4 ah = Vec16<__nv_bfloat16>::pack(a), bh = Vec16<__nv_bfloat16>::pack(b);
    stg_stream_v4(a_hi + off, ah);
    stg_stream_v4(b_hi + off, bh);
    if (a_lo != nullptr) {
      float r[8];
      Vec16<__nv_bfloat16>::unpack(ah, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = a[e] - r[e];                     // exact in fp32
      stg_stream_v4(a_lo + off, Vec16<__nv_bfloat16>::pack(r));
      Vec16<__nv_bfloat16>::unpack(bh, r);
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = b[e] - r[e];
      stg_stream_v4(b_lo + off, Vec16<__nv_bfloat16>::pack(r));
    }
  }
  if (cc * 8 < D) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      double v = acc[e];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && cc * 8 + e < D) atomicAdd(diag + cc * 8 + e, v);
    }
  }
}

// G (fp32 [D][D]) -> packed hi/lo planes of G and of G^T, zero padded to D_pad.
// thread = (chunk column cc, row r); lanes run over r.  grid = (D_pad/8, D_pad/256).
__global__ void __launch_bounds__(256) lfd_pack_g_kernel(const float* __restrict__ G, int D, int64_t D_pad,
                                                         __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo,
                                                         __nv_bfloat16* __restrict__ gt_hi,
                                                         __nv_bfloat16* __restrict__ gt_lo) {
  const int cc = blockIdx.x, c0 = cc * 8;
  const int64_t r = static_cast<int64_t>(blockIdx.y) * 256 + threadIdx.x;
  float g[8], gt[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool ok = r < D && c0 + e < D;
    g[e] = ok ? G[r * D + c0 + e] : 0.0f;               // P(G):   X[r][c] = G[r][c]
    gt[e] = ok ? G[static_cast<int64_t>(c0 + e) * D + r] : 0.0f;   // P(G^T): X[r][c] = G[c][r]
  }
  const int64_t off = (static_cast<int64_t>(cc) * D_pad + r) * 8;
  float t[8];
  uint4 h = Vec16<__nv_bfloat16>::pack(g);
  *reinterpret_cast<uint4*>(g_hi + off) = h;
  Vec16<__nv_bfloat16>::unpack(h, t);
#pragma unroll
  for (int e = 0; e < 8; ++e) t[e] = g[e] - t[e];
  *reinterpret_cast<uint4*>(g_lo + off) = Vec16<__nv_bfloat16>::pack(t);
  h = Vec16<__nv_bfloat16>::pack(gt);
  *reinterpret_cast<uint4*>(gt_hi + off) = h;
  Vec16<__nv_bfloat16>::unpack(h, t);
#pragma unroll
  for (int e = 0; e < 8; ++e) t[e] = gt[e] - t[e];
  *reinterpret_cast<uint4*>(gt_lo + off) = Vec16<__nv_bfloat16>::pack(t);
}

// cov[i] = sum_s partial[s][i], fixed order (deterministic split-K reduction); the diagonal comes
// from the fp64 CUDA-core sums
__global__ void __launch_bounds__(256) lfd_splitk_reduce_kernel(const float* __restrict__ partial, int splits,
                                                                int64_t n, int D, const double* __restrict__ diag,
                                                                float* __restrict__ cov) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t j = i / D;
    if (i - j * D == j) {
      cov[i] = static_cast<float>(diag[j]);
      continue;
    }
    float a = 0.0f;
    for (int s = 0; s < splits; ++s) a += partial[static_cast<int64_t>(s) * n + i];
    cov[i] = a;
  }
}

// loss = sum_j (1-C_jj)^2 + lambda sum_{j!=k} C_jk^2,  G = dloss/dC            (losses:51-57)
__global__ void __launch_bounds__(256) lfd_loss_kernel(const float* __restrict__ cov, int D, double inv_n, float lambda,
                                                       unsigned int* counters, double* partials,
                                                       float* __restrict__ loss_out, float* __restrict__ G) {
  __shared__ double s_red[8];
  __shared__ int s_last;
  const int64_t n = static_cast<int64_t>(D) * D;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  double acc = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int j = static_cast<int>(i / D), k = static_cast<int>(i - static_cast<int64_t>(j) * D);
    const double c = static_cast<double>(cov[i]) * inv_n;
    if (j == k) {
      acc += (1.0 - c) * (1.0 - c);
      G[i] = static_cast<float>(-2.0 * (1.0 - c));
    } else {
      acc += static_cast<double>(lambda) * c * c;
      G[i] = static_cast<float>(2.0 * static_cast<double>(lambda) * c);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    partials[blockIdx.x] = t;
    __threadfence();
    s_last = (atomicAdd(&counters[0], 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {                                        // fixed-order tree over the per-CTA partials
    __threadfence();
    double t = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(&partials[b]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < 8; ++w) tot += s_red[w];
      *loss_out = static_cast<float>(tot);
      counters[0] = 0;
      __threadfence();
    }
  }
}

// Small-batch path (B < 32, natural row order): sum_b dz~*z~ per (t,d) for both tensors -> bn[2][TD] fp32
// (fp64 accumulation).  The other batch-norm moment, sum_b dz~ = sum_k (sum_b z~[b,t,k]) G[.,k] / N, vanishes
// identically because z~ has zero batch mean, and is not computed.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) lfd_bn_reduce_kernel(const T* __restrict__ za, const T* __restrict__ zb,
                                                            const float* __restrict__ dza, const float* __restrict__ dzb,
                                                            const float* __restrict__ tables, int B, int64_t TD,
                                                            int bsplit, float* __restrict__ bn) {
  constexpr int N = VEC ? Vec16<T>::N : 1;
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v * N >= TD) return;
  const int which = blockIdx.y;
  const T* z = which == 0 ? za : zb;
  const float* dz = which == 0 ? dza : dzb;
  float sc[N], sh[N];
#pragma unroll
  for (int e = 0; e < N; ++e) {
    sc[e] = tables[which * 2 * TD + v * N + e];
    sh[e] = tables[which * 2 * TD + TD + v * N + e];
  }
  const int bchunk = (B + bsplit - 1) / bsplit;
  const int b0 = blockIdx.z * bchunk, b1 = min(B, b0 + bchunk);
  double s2[N];
#pragma unroll
  for (int e = 0; e < N; ++e) s2[e] = 0.0;
  for (int b = b0; b < b1; ++b) {
    float x[N], g[N];
    const int64_t off = static_cast<int64_t>(b) * TD + v * N;
    if (VEC) {
      Vec16<T>::unpack(ldg_stream_v4(z + off), x);
#pragma unroll
      for (int q = 0; q < N / 4; ++q) Vec16<float>::unpack(ldg_stream_v4(dz + off + 4 * q), g + 4 * q);
    } else {
      x[0] = Vec16<T>::load1(z + off);
      g[0] = dz[off];
    }
#pragma unroll
    for (int e = 0; e < N; ++e) {
      const float zt = fmaf(x[e], sc[e], sh[e]);
      s2[e] += static_cast<double>(g[e]) * static_cast<double>(zt);
    }
  }
  float* o2 = bn + static_cast<int64_t>(which) * TD + v * N;
#pragma unroll
  for (int e = 0; e < N; ++e) {
    if (bsplit == 1) o2[e] = static_cast<float>(s2[e]);
    else atomicAdd(o2 + e, static_cast<float>(s2[e]));
  }
}

// dx = (dz~ - mean_b dz~ - z~ * mean_b(dz~ z~)) * rstd * upstream      (batch-norm backward; mean_b dz~ == 0)
// Per (t,d) the expression is affine in (dz~, x):  dx = A*dz~ + Bx*x + C  with
//   A = rstd*up,  Bx = -rstd^2 * m2 * up,  C = -shift*m2 * rstd * up            (z~ = x*rstd + shift)
// m2 = (sum over the `parts` partials bn[which][t][p][d], fixed order) / n_batch.
// so a thread owns one 16-byte vector of the (t,d) plane, computes its coefficients once and streams
// over the batch (same shape as the stats kernel).  grid = (ceil(TD/N/256), 2, bsplit).
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) lfd_bn_finalize_kernel(const T* __restrict__ za, const T* __restrict__ zb,
                                                              const float* __restrict__ dza,
                                                              const float* __restrict__ dzb,
                                                              const float* __restrict__ tables,
                                                              const float* __restrict__ bn, int parts, int D,
                                                              double inv_nb, const float* __restrict__ grad_scale, int B,
                                                              int64_t TD, int bsplit, T* __restrict__ oa,
                                                              T* __restrict__ ob) {
  constexpr int N = VEC ? Vec16<T>::N : 1;
  const int64_t v = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v * N >= TD) return;
  const int which = blockIdx.y;
  const T* z = which == 0 ? za : zb;
  const float* dz = which == 0 ? dza : dzb;
  T* out = which == 0 ? oa : ob;
  const float up = grad_scale ? __ldg(grad_scale) : 1.0f;
  float cA[N], cB[N], cC[N];
#pragma unroll
  for (int e = 0; e < N; ++e) {
    const int64_t td = v * N + e;
    const float sc = tables[which * 2 * TD + td], sh = tables[which * 2 * TD + TD + td];
    const int64_t tt = td / D, dd = td - tt * D;
    const float* bp = bn + (static_cast<int64_t>(which) * (TD / D) + tt) * parts * D + dd;
    float acc = 0.0f;
    for (int pp = 0; pp < parts; ++pp) acc += bp[static_cast<int64_t>(pp) * D];
    const float m2 = static_cast<float>(static_cast<double>(acc) * inv_nb);
    cA[e] = sc * up;
    cB[e] = -sc * sc * m2 * up;
    cC[e] = -(sh * m2) * sc * up;
  }
  const int bchunk = (B + bsplit - 1) / bsplit;
  const int b0 = blockIdx.z * bchunk, b1 = min(B, b0 + bchunk);
  constexpr int U = 2;
  int b = b0;
  for (; b + U <= b1; b += U) {
    float x[U][N], g[U][N];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t off = static_cast<int64_t>(b + u) * TD + v * N;
      if (VEC) {
        Vec16<T>::unpack(ldg_stream_v4(z + off), x[u]);
#pragma unroll
        for (int q = 0; q < N / 4; ++q) Vec16<float>::unpack(ldg_stream_v4(dz + off + 4 * q), g[u] + 4 * q);
      } else {
        x[u][0] = Vec16<T>::load1(z + off);
        g[u][0] = dz[off];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t off = static_cast<int64_t>(b + u) * TD + v * N;
      float o[N];
#pragma unroll
      for (int e = 0; e < N; ++e) o[e] = fmaf(cA[e], g[u][e], fmaf(cB[e], x[u][e], cC[e]));
      if (VEC) stg_stream_v4(out + off, Vec16<T>::pack(o));
      else Vec16<T>::store1(out + off, o[0]);
    }
  }
  for (; b < b1; ++b) {
    const int64_t off = static_cast<int64_t>(b) * TD + v * N;
    float x[N], g[N], o[N];
    if (VEC) {
      Vec16<T>::unpack(ldg_stream_v4(z + off), x);
#pragma unroll
      for (int q = 0; q < N / 4; ++q) Vec16<float>::unpack(ldg_stream_v4(dz + off + 4 * q), g + 4 * q);
    } else {
      x[0] = Vec16<T>::load1(z + off);
      g[0] = dz[off];
    }
#pragma unroll
    for (int e = 0; e < N; ++e) o[e] = fmaf(cA[e], g[e], fmaf(cB[e], x[e], cC[e]));
    if (VEC) stg_stream_v4(out + off, Vec16<T>::pack(o));
    else Vec16<T>::store1(out + off, o[0]);
  }
}

int check_common(const char* what, const void* za, const void* zb, int dtype, int64_t B, int64_t T, int64_t D) {
  FDDM_CHECK_ARG(za && zb, "%s: null input", what);
  FDDM_CHECK_ARG(dtype_valid(dtype), "%s: bad dtype %d", what, dtype);
  FDDM_CHECK_ARG(B > 0 && T > 0 && D > 0, "%s: non-positive size", what);
  FDDM_CHECK_ARG(B < (1ll << 31) && T * D < (1ll << 40) && B * T < (1ll << 31), "%s: size too large", what);
  return FDDM_OK;
}

bool vec_ok(const void* a, const void* b, int dtype, int64_t TD) {
  const int n = dtype == FDDM_F32 ? 4 : 8;
  return TD % n == 0 && reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(b) % 16 == 0;
}

int pick_bsplit(int64_t nthreads_plane, int64_t B) {
  // fill ~2 waves of 148 SMs x 8 CTAs of 256 threads when the (t,d) plane alone is too small
  const int64_t want = 148ll * 2048;
  int64_t s = (want + nthreads_plane - 1) / nthreads_plane;
  s = std::min<int64_t>(s, std::max<int64_t>(1, B / 8));
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(s, 64)));
}

int launch_tables(const double* sums, int64_t TD, double n_batch, float eps, float* tables, cudaStream_t stream) {
  const int64_t n = 2 * TD;
  KernelScope ks("lfd_tables_kernel", stream);
  lfd_tables_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(sums, TD, n_batch, eps, tables);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

template <typename T>
int launch_stats(const void* za, const void* zb, int dtype, int64_t B, int64_t Tn, int64_t D, double* sums,
                 cudaStream_t stream) {
  const int64_t TD = Tn * D;
  if (vec_ok(za, zb, dtype, D) && Tn < 65536) {
    KernelScope ks("lfd_stats_kernel", stream);
    constexpr int N = Vec16<T>::N;
    if (B <= 128) {
      dim3 grid(static_cast<unsigned>((D + 32 * N - 1) / (32 * N)), static_cast<unsigned>(Tn), 2);
      lfd_stats_rows_kernel<T, 32, 8, 8><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb),
                                                                   static_cast<int>(B), static_cast<int>(Tn),
                                                                   static_cast<int>(D), sums);
    } else {
      dim3 grid(static_cast<unsigned>((D + 64 * N - 1) / (64 * N)), static_cast<unsigned>(Tn), 2);
      lfd_stats_rows_kernel<T, 64, 4, 4><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb),
                                                                   static_cast<int>(B), static_cast<int>(Tn),
                                                                   static_cast<int>(D), sums);
    }
    FDDM_LAUNCH_OK();
    return FDDM_OK;
  }
  const bool vec = vec_ok(za, zb, dtype, TD);
  const int n = vec ? Vec16<T>::N : 1;
  const int64_t nthr = (TD + n - 1) / n;
  const int bsplit = pick_bsplit(nthr, B);
  KernelScope ks("lfd_stats_kernel", stream);
  if (bsplit > 1) FDDM_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 4 * TD, stream));
  dim3 grid(static_cast<unsigned>((nthr + 255) / 256), 2, bsplit);
  if (vec)
    lfd_stats_kernel<T, true><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb),
                                                        static_cast<int>(B), TD, bsplit, sums);
  else
    lfd_stats_kernel<T, false><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb),
                                                         static_cast<int>(B), TD, bsplit, sums);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

template <typename T>
int launch_bn(const void* za, const void* zb, int dtype, int64_t B, int64_t TD, const float* dza, const float* dzb,
              const float* tables, float* bn, cudaStream_t stream) {
  const bool vec = vec_ok(za, zb, dtype, TD) && TD % 8 == 0;
  const int n = vec ? Vec16<T>::N : 1;
  const int64_t nthr = (TD + n - 1) / n;
  const int bsplit = pick_bsplit(nthr, B);
  KernelScope ks("lfd_bn_reduce_kernel", stream);
  if (bsplit > 1) FDDM_CUDA_OK(cudaMemsetAsync(bn, 0, sizeof(float) * 2 * TD, stream));
  dim3 grid(static_cast<unsigned>((nthr + 255) / 256), 2, bsplit);
  if (vec)
    lfd_bn_reduce_kernel<T, true><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb), dza,
                                                            dzb, tables, static_cast<int>(B), TD, bsplit, bn);
  else
    lfd_bn_reduce_kernel<T, false><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb), dza,
                                                             dzb, tables, static_cast<int>(B), TD, bsplit, bn);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

template <typename T>
int launch_finalize(const void* za, const void* zb, int dtype, int64_t B, int64_t TD, const float* dza,
                    const float* dzb, const float* tables, const float* bn, int parts, int D, double inv_nb,
                    const float* grad_scale, void* oa, void* ob, cudaStream_t stream) {
  const bool vec = vec_ok(za, zb, dtype, TD) && TD % 8 == 0 && reinterpret_cast<uintptr_t>(oa) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(ob) % 16 == 0;
  const int n = vec ? Vec16<T>::N : 1;
  const int64_t nthr = (TD + n - 1) / n;
  const int bsplit = pick_bsplit(nthr, B);
  dim3 grid(static_cast<unsigned>((nthr + 255) / 256), 2, bsplit);
  KernelScope ks("lfd_bn_finalize_kernel", stream);
  if (vec)
    lfd_bn_finalize_kernel<T, true><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb), dza,
                                                              dzb, tables, bn, parts, D, inv_nb, grad_scale,
                                                              static_cast<int>(B), TD, bsplit, static_cast<T*>(oa),
                                                              static_cast<T*>(ob));
  else
    lfd_bn_finalize_kernel<T, false><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb),
                                                               dza, dzb, tables, bn, parts, D, inv_nb, grad_scale,
                                                               static_cast<int>(B), TD, bsplit, static_cast<T*>(oa),
                                                               static_cast<T*>(ob));
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

#define FDDM_DISPATCH_DT(dtype_, FN, ...)                                  \
  ((dtype_) == FDDM_F32 ? FN<float>(__VA_ARGS__)                           \
                        : ((dtype_) == FDDM_BF16 ? FN<__nv_bfloat16>(__VA_ARGS__) : FN<__half>(__VA_ARGS__)))

// number of split-K slices for the forward contraction: fill the SMs, >= 2 k-blocks per slice
int pick_splits(int64_t tiles, int64_t rows) {
  const int sms = num_sms();
  int64_t s = std::max<int64_t>(1, sms / tiles);             // tiles*splits <= SM count: exactly one wave
  s = std::min<int64_t>(s, std::max<int64_t>(1, rows / 64));
  s = std::min<int64_t>(s, LfdWorkspace::kMaxSplits);
  return static_cast<int>(s);
}


int check_gemm_shape(const char* what, const void* za, const void* zb, int64_t D) {
  if (D % 8 != 0) {
    set_error("%s: D=%lld must be a multiple of 8 for the tensor-core contraction", what, (long long)D);
    return FDDM_EUNSUPPORTED;
  }
  FDDM_CHECK_ARG(reinterpret_cast<uintptr_t>(za) % 16 == 0 && reinterpret_cast<uintptr_t>(zb) % 16 == 0,
                 "%s: inputs must be 16-byte aligned", what);
  return FDDM_OK;
}

// tables + the one-pass standardise / split / pack (+ fp64 diagonal)
int standardise_and_pack(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D,
                         const double* sums, double n_batch_global, float eps, uint8_t* ws, const LfdWorkspace& lay,
                         int terms, cudaStream_t stream) {
  float* tables = reinterpret_cast<float*>(ws + lay.off_tables);
  double* diag = reinterpret_cast<double*>(ws + lay.off_diag);
  const LfdRows rm(B, T);
  const int64_t TD = T * D, rows = B * T, Rp = pack_pad(rm.rows_packed), Dp = pack_pad(D);
  if (int rc = launch_tables(sums, TD, n_batch_global, eps, tables, stream)) return rc;
  FDDM_CUDA_OK(cudaMemsetAsync(diag, 0, sizeof(double) * D, stream));
  __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(ws + lay.off_pack);
  const size_t pe = lay.plane_bytes / 2;
  __nv_bfloat16 *a_hi = pl, *a_lo = terms == 2 ? pl + pe : nullptr, *b_hi = pl + 2 * pe,
                *b_lo = terms == 2 ? pl + 3 * pe : nullptr;
  KernelScope ks("lfd_pack_kernel", stream);
  if (rm.tb_major) {
    FDDM_CHECK_ARG(T < 65536, "lfd: T too large for the tb-major pack grid");
    dim3 gtb(static_cast<unsigned>(Dp / 64), static_cast<unsigned>(T));
#define FDDM_PACK_TB(TY)                                                                                              \
  lfd_pack_tb_kernel<TY><<<gtb, 256, 0, stream>>>(static_cast<const TY*>(z_a), static_cast<const TY*>(z_b), tables,  \
                                                   rm.B, rm.Bp, rm.T, static_cast<int>(D), Rp, a_hi, a_lo, b_hi, b_lo, diag)
    if (dtype == FDDM_F32) FDDM_PACK_TB(float);
    else if (dtype == FDDM_BF16) FDDM_PACK_TB(__nv_bfloat16);
    else FDDM_PACK_TB(__half);
#undef FDDM_PACK_TB
    FDDM_LAUNCH_OK();
    return FDDM_OK;
  }
  dim3 grid(static_cast<unsigned>(Dp / 64), static_cast<unsigned>(Rp / 256));
  if (dtype == FDDM_F32)
    lfd_pack_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(z_a), static_cast<const float*>(z_b),
                                                     tables, rows, static_cast<int>(T), static_cast<int>(D), Rp, a_hi,
                                                     a_lo, b_hi, b_lo, diag);
  else if (dtype == FDDM_BF16)
    lfd_pack_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(z_a),
                                                             static_cast<const __nv_bfloat16*>(z_b), tables, rows,
                                                             static_cast<int>(T), static_cast<int>(D), Rp, a_hi, a_lo,
                                                             b_hi, b_lo, diag);
  else
    lfd_pack_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(z_a), static_cast<const __half*>(z_b),
                                                      tables, rows, static_cast<int>(T), static_cast<int>(D), Rp, a_hi,
                                                      a_lo, b_hi, b_lo, diag);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

void planes_of(uint8_t* ws, const LfdWorkspace& lay, int64_t B, int64_t T, int64_t D, int terms, int mn_is_col,
               PackedOperand& A, PackedOperand& Bo) {
  __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(ws + lay.off_pack);
  const size_t pe = lay.plane_bytes / 2;
  A.hi = pl; A.lo = terms == 2 ? pl + pe : nullptr;
  Bo.hi = pl + 2 * pe; Bo.lo = terms == 2 ? pl + 3 * pe : nullptr;
  A.R_pad = Bo.R_pad = pack_pad(LfdRows(B, T).rows_packed);
  A.C_pad = Bo.C_pad = pack_pad(D);
  A.mn_is_col = Bo.mn_is_col = mn_is_col;
}

}  // namespace
}  // namespace fddm

extern "C" {

size_t fddm_lfd_workspace_bytes(int64_t B, int64_t T, int64_t D) {
  if (B <= 0 || T <= 0 || D <= 0) return 0;
  return fddm::LfdWorkspace(B, T, D).total;
}

int fddm_lfd_stats(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D, double* sums,
                   fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_common("lfd_stats", z_a, z_b, dtype, B, T, D)) return rc;
  FDDM_CHECK_ARG(sums, "lfd_stats: null sums");
  return FDDM_DISPATCH_DT(dtype, launch_stats, z_a, z_b, dtype, B, T, D, sums, stream);
}

int fddm_lfd_xcov(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D, const double* sums,
                  double n_batch_global, float eps, void* workspace, float* cov, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_common("lfd_xcov", z_a, z_b, dtype, B, T, D)) return rc;
  FDDM_CHECK_ARG(sums && workspace && cov, "lfd_xcov: null pointer argument");
  FDDM_CHECK_ARG(n_batch_global >= static_cast<double>(B), "lfd_xcov: n_batch_global < B");
  if (int rc = check_gemm_shape("lfd_xcov", z_a, z_b, D)) return rc;
  const LfdWorkspace lay(B, T, D);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* partial = reinterpret_cast<float*>(ws + lay.off_splitk);
  const int64_t rows = LfdRows(B, T).rows_packed;       // contraction length incl. the zero rows of batch padding
  const int terms = (dtype == FDDM_BF16) ? 1 : 2;       // bf16 inputs: the reference's matmul operands are bf16 too
  if (int rc = standardise_and_pack(z_a, z_b, dtype, B, T, D, sums, n_batch_global, eps, ws, lay, terms, stream))
    return rc;
  PackedOperand A{}, Bo{};
  planes_of(ws, lay, B, T, D, terms, 1, A, Bo);
  const int64_t tiles = ((D + 127) / 128) * ((D + 255) / 256);
  const int splits = pick_splits(tiles, rows);
  if (int rc = umma_gemm(A, Bo, D, D, rows, splits, terms, 1.0f, partial, D, D * D, stream)) return rc;
  const int64_t n = D * D;
  KernelScope ks("lfd_splitk_reduce_kernel", stream);
  lfd_splitk_reduce_kernel<<<static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, num_sms() * 8)), 256, 0,
                             stream>>>(partial, splits, n, static_cast<int>(D),
                                       reinterpret_cast<const double*>(ws + lay.off_diag), cov);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int fddm_lfd_loss(const float* cov, int64_t D, double n_rows_global, float lambda_offdiag, void* workspace,
                  float* loss_out, float* G, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(cov && workspace && loss_out && G, "lfd_loss: null pointer argument");
  FDDM_CHECK_ARG(D > 0 && D < 46341 && n_rows_global > 0.0, "lfd_loss: bad size");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int grid = static_cast<int>(std::min<int64_t>((D * D + 255) / 256, std::min<int64_t>(num_sms() * 4,
                                                                              LfdWorkspace::kMaxPartials)));
  KernelScope ks("lfd_loss_kernel", stream);
  lfd_loss_kernel<<<grid, 256, 0, stream>>>(cov, static_cast<int>(D), 1.0 / n_rows_global, lambda_offdiag,
                                            reinterpret_cast<unsigned int*>(ws),
                                            reinterpret_cast<double*>(ws + LfdWorkspace::kCounters), loss_out, G);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int64_t fddm_lfd_bn_parts(int64_t B, int64_t T, int64_t D) {
  (void)D;
  if (B <= 0 || T <= 0) return 0;
  return fddm::LfdRows(B, T).parts();
}

int fddm_lfd_backward(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D, const double* sums,
                      double n_batch_global, float eps, const float* G, double n_rows_global, const float* grad_scale,
                      void* workspace, float* bn_sums, int64_t bn_parts, int phase, void* dz_a, void* dz_b,
                      fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_common("lfd_backward", z_a, z_b, dtype, B, T, D)) return rc;
  FDDM_CHECK_ARG(sums && G && workspace && bn_sums && dz_a && dz_b, "lfd_backward: null pointer argument");
  const bool planes_valid = (phase & FDDM_LFD_PLANES_VALID) != 0;
  phase &= ~FDDM_LFD_PLANES_VALID;
  FDDM_CHECK_ARG(phase == 0 || phase == 1, "lfd_backward: phase must be 0 or 1");
  const LfdRows rm(B, T);
  FDDM_CHECK_ARG(phase == 0 ? bn_parts == rm.parts() : (bn_parts == rm.parts() || bn_parts == 1),
                 "lfd_backward: bn_parts must be fddm_lfd_bn_parts(B,T,D) (phase 1 also accepts 1: partials already summed)");
  if (int rc = check_gemm_shape("lfd_backward", z_a, z_b, D)) return rc;
  const LfdWorkspace lay(B, T, D);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* tables = reinterpret_cast<float*>(ws + lay.off_tables);
  float* dza = reinterpret_cast<float*>(ws + lay.off_dza);
  float* dzb = reinterpret_cast<float*>(ws + lay.off_dzb);
  const int64_t TD = T * D, rows = B * T;
  if (phase == 0) {
    // the forward's tables / planes are reused only when the caller vouches that this workspace still
    // holds them; the backward always needs the residual planes (G is fp32: hi + residual for all dtypes)
    const bool have_lo = (dtype != FDDM_BF16);
    if (!planes_valid || !have_lo)
      if (int rc = standardise_and_pack(z_a, z_b, dtype, B, T, D, sums, n_batch_global, eps, ws, lay, 2, stream))
        return rc;
    const int64_t Dp = pack_pad(D);
    __nv_bfloat16* gp = reinterpret_cast<__nv_bfloat16*>(ws + lay.off_gpack);
    const size_t ge = lay.gplane_bytes / 2;
    {
      KernelScope ks("lfd_pack_g_kernel", stream);
      lfd_pack_g_kernel<<<dim3(static_cast<unsigned>(Dp / 8), static_cast<unsigned>(Dp / 256)), 256, 0, stream>>>(
          G, static_cast<int>(D), Dp, gp, gp + ge, gp + 2 * ge, gp + 3 * ge);
      FDDM_LAUNCH_OK();
    }
    const float alpha = static_cast<float>(1.0 / n_rows_global);
    PackedOperand Za{}, Zb{}, Gk{}, Gt{};
    planes_of(ws, lay, B, T, D, 2, 0, Za, Zb);
    Gk.hi = gp; Gk.lo = gp + ge; Gk.R_pad = Dp; Gk.C_pad = Dp; Gk.mn_is_col = 0;            // B(n=j, k) = G[j][k]
    Gt = Gk; Gt.hi = gp + 2 * ge; Gt.lo = gp + 3 * ge;                                      // B(n=k, j) = G^T[k][j]
    if (rm.tb_major) {
      // persistent contraction whose epilogue also emits the batch sums of dz~ * z~ as [T][Bp/32][D] partials
      float* bn_a = bn_sums;
      float* bn_b = bn_sums + static_cast<int64_t>(T) * rm.parts() * D;
      if (int rc = umma_bwd_gemm(Zb, Gk, Za, rm, D, alpha, dza, bn_a, stream)) return rc;
      return umma_bwd_gemm(Za, Gt, Zb, rm, D, alpha, dzb, bn_b, stream);
    }
    // dza~[r][j] = (1/N) sum_k zb~[r][k] G[j][k]          (oracle: B2 @ G.T / N)
    if (int rc = umma_gemm(Zb, Gk, rows, D, D, 1, 2, alpha, dza, D, 0, stream)) return rc;
    // dzb~[r][k] = (1/N) sum_j za~[r][j] G[j][k]          (oracle: A2 @ G / N)
    if (int rc = umma_gemm(Za, Gt, rows, D, D, 1, 2, alpha, dzb, D, 0, stream)) return rc;
    return FDDM_DISPATCH_DT(dtype, launch_bn, z_a, z_b, dtype, B, TD, dza, dzb, tables, bn_sums, stream);
  }
  return FDDM_DISPATCH_DT(dtype, launch_finalize, z_a, z_b, dtype, B, TD, dza, dzb, tables, bn_sums,
                          static_cast<int>(bn_parts), static_cast<int>(D), 1.0 / n_batch_global, grad_scale, dz_a, dz_b,
                          stream);
}

}  // extern "C"
