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

// The diagonal of cov on the CUDA cores with fp64 accumulation: diag[j] += sum_rows za~[r,j] zb~[r,j].
// Why: the tensor-core accumulator adds with truncation, harmless for the zero-mean off-diagonal
// sums but a systematic ~1e-6 relative bias on the all-positive diagonal sums when z_a and z_b are
// correlated -- and the loss is dominated by sum_j (1-C_jj)^2.  One more read of the inputs (s bytes
// per element), D results.  grid = (ceil(D/N/64), row slices); rows are strided over blockIdx.y.
template <typename T>
__global__ void __launch_bounds__(256) lfd_diag_kernel(const T* __restrict__ za, const T* __restrict__ zb,
                                                       const float* __restrict__ tables, int64_t rows, int Tn, int D,
                                                       double* __restrict__ diag) {
  constexpr int N = Vec16<T>::N;
  __shared__ double s_acc[4][64][N];
  const int vx = threadIdx.x & 63, ry = threadIdx.x >> 6;       // 64 d-vectors x 4 row lanes per CTA
  const int v = blockIdx.x * 64 + vx;
  const int64_t TD = static_cast<int64_t>(Tn) * D;
  double acc[N];
#pragma unroll
  for (int e = 0; e < N; ++e) acc[e] = 0.0;
  if (v * N < D) {
    for (int64_t r = static_cast<int64_t>(blockIdx.y) * 4 + ry; r < rows; r += static_cast<int64_t>(gridDim.y) * 4) {
      const int64_t so = (r % Tn) * D + v * N;
      float a[N], b[N];
      Vec16<T>::unpack(ldg_stream_v4(za + r * D + v * N), a);
      Vec16<T>::unpack(ldg_stream_v4(zb + r * D + v * N), b);
#pragma unroll
      for (int e = 0; e < N; ++e) {
        const float at = fmaf(a[e], tables[so + e], tables[TD + so + e]);
        const float bt = fmaf(b[e], tables[2 * TD + so + e], tables[3 * TD + so + e]);
        acc[e] += static_cast<double>(at) * static_cast<double>(bt);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < N; ++e) s_acc[ry][vx][e] = acc[e];
  __syncthreads();
  if (ry == 0 && v * N < D) {
#pragma unroll
    for (int e = 0; e < N; ++e)
      atomicAdd(diag + v * N + e, (s_acc[0][vx][e] + s_acc[1][vx][e]) + (s_acc[2][vx][e] + s_acc[3][vx][e]));
  }
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
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(&partials[b]);   // fixed order
    *loss_out = static_cast<float>(t);
    counters[0] = 0;
    __threadfence();
  }
}

// sum_b dz~ and sum_b dz~*z~ per (t,d) for both tensors -> bn[2][2][TD]  (same shape as stats)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) lfd_bn_reduce_kernel(const T* __restrict__ za, const T* __restrict__ zb,
                                                            const float* __restrict__ dza, const float* __restrict__ dzb,
                                                            const float* __restrict__ tables, int B, int64_t TD,
                                                            int bsplit, double* __restrict__ bn) {
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
  double s1[N], s2[N];
#pragma unroll
  for (int e = 0; e < N; ++e) { s1[e] = 0.0; s2[e] = 0.0; }
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
      s1[e] += static_cast<double>(g[e]);
      s2[e] += static_cast<double>(g[e]) * static_cast<double>(zt);
    }
  }
  double* o1 = bn + static_cast<int64_t>(which) * 2 * TD + v * N;
  double* o2 = o1 + TD;
#pragma unroll
  for (int e = 0; e < N; ++e) {
    if (bsplit == 1) { o1[e] = s1[e]; o2[e] = s2[e]; }
    else { atomicAdd(o1 + e, s1[e]); atomicAdd(o2 + e, s2[e]); }
  }
}

// dx = (dz~ - mean_b dz~ - z~ * mean_b(dz~ z~)) * rstd * upstream      (batch-norm backward)
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) lfd_bn_finalize_kernel(const T* __restrict__ za, const T* __restrict__ zb,
                                                              const float* __restrict__ dza,
                                                              const float* __restrict__ dzb,
                                                              const float* __restrict__ tables,
                                                              const double* __restrict__ bn, double inv_nb,
                                                              const float* __restrict__ grad_scale, int B, int64_t TD,
                                                              T* __restrict__ oa, T* __restrict__ ob) {
  constexpr int N = VEC ? Vec16<T>::N : 1;
  const int which = blockIdx.y;
  const T* z = which == 0 ? za : zb;
  const float* dz = which == 0 ? dza : dzb;
  T* out = which == 0 ? oa : ob;
  const float up = grad_scale ? __ldg(grad_scale) : 1.0f;
  const int64_t nvec_plane = (TD + N - 1) / N;
  const int64_t total = static_cast<int64_t>(B) * nvec_plane;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / nvec_plane, v = i - b * nvec_plane;
    const int64_t td = v * N, off = b * TD + td;
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
    for (int e = 0; e < N; ++e) {
      const float sc = tables[which * 2 * TD + td + e], sh = tables[which * 2 * TD + TD + td + e];
      const float m1 = static_cast<float>(bn[which * 2 * TD + td + e] * inv_nb);
      const float m2 = static_cast<float>(bn[which * 2 * TD + TD + td + e] * inv_nb);
      const float zt = fmaf(x[e], sc, sh);
      o[e] = (g[e] - m1 - zt * m2) * sc * up;
    }
    if (VEC) stg_stream_v4(out + off, Vec16<T>::pack(o));
    else Vec16<T>::store1(out + off, o[0]);
  }
}

struct LfdArgs {
  const void* za; const void* zb; int dtype; int64_t B, T, D;
};

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
  lfd_tables_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(sums, TD, n_batch, eps, tables);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

template <typename T>
int launch_stats(const void* za, const void* zb, int dtype, int64_t B, int64_t TD, double* sums, cudaStream_t stream) {
  const bool vec = vec_ok(za, zb, dtype, TD);
  const int n = vec ? Vec16<T>::N : 1;
  const int64_t nthr = (TD + n - 1) / n;
  const int bsplit = pick_bsplit(nthr, B);
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
              const float* tables, double* bn, cudaStream_t stream) {
  const bool vec = vec_ok(za, zb, dtype, TD) && TD % 8 == 0;
  const int n = vec ? Vec16<T>::N : 1;
  const int64_t nthr = (TD + n - 1) / n;
  const int bsplit = pick_bsplit(nthr, B);
  if (bsplit > 1) FDDM_CUDA_OK(cudaMemsetAsync(bn, 0, sizeof(double) * 4 * TD, stream));
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
                    const float* dzb, const float* tables, const double* bn, double inv_nb, const float* grad_scale,
                    void* oa, void* ob, cudaStream_t stream) {
  const bool vec = vec_ok(za, zb, dtype, TD) && TD % 8 == 0 && reinterpret_cast<uintptr_t>(oa) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(ob) % 16 == 0;
  dim3 grid(static_cast<unsigned>(num_sms() * 8), 2, 1);
  if (vec)
    lfd_bn_finalize_kernel<T, true><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb), dza,
                                                              dzb, tables, bn, inv_nb, grad_scale, static_cast<int>(B),
                                                              TD, static_cast<T*>(oa), static_cast<T*>(ob));
  else
    lfd_bn_finalize_kernel<T, false><<<grid, 256, 0, stream>>>(static_cast<const T*>(za), static_cast<const T*>(zb),
                                                               dza, dzb, tables, bn, inv_nb, grad_scale,
                                                               static_cast<int>(B), TD, static_cast<T*>(oa),
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
  int64_t s = std::max<int64_t>(1, (sms + tiles - 1) / tiles);
  s = std::min<int64_t>(s, std::max<int64_t>(1, rows / 64));
  s = std::min<int64_t>(s, LfdWorkspace::kMaxSplits);
  return static_cast<int>(s);
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
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_common("lfd_stats", z_a, z_b, dtype, B, T, D)) return rc;
  FDDM_CHECK_ARG(sums, "lfd_stats: null sums");
  return FDDM_DISPATCH_DT(dtype, launch_stats, z_a, z_b, dtype, B, T * D, sums, stream);
}

int fddm_lfd_xcov(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D, const double* sums,
                  double n_batch_global, float eps, void* workspace, float* cov, fddm_stream_t stream_) {
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_common("lfd_xcov", z_a, z_b, dtype, B, T, D)) return rc;
  FDDM_CHECK_ARG(sums && workspace && cov, "lfd_xcov: null pointer argument");
  FDDM_CHECK_ARG(reinterpret_cast<uintptr_t>(z_a) % 16 == 0 && reinterpret_cast<uintptr_t>(z_b) % 16 == 0,
                 "lfd_xcov: inputs must be 16-byte aligned");
  FDDM_CHECK_ARG(n_batch_global >= static_cast<double>(B), "lfd_xcov: n_batch_global < B");
  if (D % 8 != 0) {
    set_error("lfd_xcov: D=%lld must be a multiple of 8 for the tensor-core contraction", (long long)D);
    return FDDM_EUNSUPPORTED;
  }
  const LfdWorkspace lay(B, T, D);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* tables = reinterpret_cast<float*>(ws + lay.off_tables);
  float* partial = reinterpret_cast<float*>(ws + lay.off_splitk);
  const int64_t TD = T * D, rows = B * T;
  if (int rc = launch_tables(sums, TD, n_batch_global, eps, tables, stream)) return rc;

  UmmaOperand A{}, Bo{};
  A.ptr = z_a; A.dtype = dtype; A.ld = D; A.nrows = rows; A.ncols = D; A.mn_is_col = 1;
  A.scale = tables; A.shift = tables + TD; A.T = static_cast<int>(T); A.stat_ld = D;
  Bo = A;
  Bo.ptr = z_b; Bo.scale = tables + 2 * TD; Bo.shift = tables + 3 * TD;
  const int64_t tiles = ((D + 127) / 128) * ((D + 255) / 256);
  const int splits = pick_splits(tiles, rows);
  const int terms = (dtype == FDDM_BF16) ? 1 : 2;
  if (int rc = umma_gemm(A, Bo, D, D, rows, splits, terms, 1.0f, partial, D, D * D, stream)) return rc;
  double* diag = reinterpret_cast<double*>(ws + lay.off_diag);
  FDDM_CUDA_OK(cudaMemsetAsync(diag, 0, sizeof(double) * D, stream));
  {
    const int nvec = static_cast<int>(D / (dtype == FDDM_F32 ? 4 : 8));
    const int gx = (nvec + 63) / 64;
    const int gy = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((rows + 3) / 4, (num_sms() * 8 + gx - 1) / gx)));
    dim3 grid(gx, gy, 1);
    if (dtype == FDDM_F32)
      lfd_diag_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(z_a), static_cast<const float*>(z_b),
                                                       tables, rows, static_cast<int>(T), static_cast<int>(D), diag);
    else if (dtype == FDDM_BF16)
      lfd_diag_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(z_a),
                                                               static_cast<const __nv_bfloat16*>(z_b), tables, rows,
                                                               static_cast<int>(T), static_cast<int>(D), diag);
    else
      lfd_diag_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(z_a), static_cast<const __half*>(z_b),
                                                        tables, rows, static_cast<int>(T), static_cast<int>(D), diag);
    FDDM_LAUNCH_OK();
  }
  const int64_t n = D * D;
  lfd_splitk_reduce_kernel<<<static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, num_sms() * 8)), 256, 0,
                             stream>>>(partial, splits, n, static_cast<int>(D), diag, cov);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int fddm_lfd_loss(const float* cov, int64_t D, double n_rows_global, float lambda_offdiag, void* workspace,
                  float* loss_out, float* G, fddm_stream_t stream_) {
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(cov && workspace && loss_out && G, "lfd_loss: null pointer argument");
  FDDM_CHECK_ARG(D > 0 && D < 46341 && n_rows_global > 0.0, "lfd_loss: bad size");
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int grid = static_cast<int>(std::min<int64_t>((D * D + 255) / 256, std::min<int64_t>(num_sms() * 4,
                                                                              LfdWorkspace::kMaxPartials)));
  lfd_loss_kernel<<<grid, 256, 0, stream>>>(cov, static_cast<int>(D), 1.0 / n_rows_global, lambda_offdiag,
                                            reinterpret_cast<unsigned int*>(ws),
                                            reinterpret_cast<double*>(ws + LfdWorkspace::kCounters), loss_out, G);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int fddm_lfd_backward(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D, const double* sums,
                      double n_batch_global, float eps, const float* G, double n_rows_global, const float* grad_scale,
                      void* workspace, double* bn_sums, int phase, void* dz_a, void* dz_b, fddm_stream_t stream_) {
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (int rc = check_common("lfd_backward", z_a, z_b, dtype, B, T, D)) return rc;
  FDDM_CHECK_ARG(sums && G && workspace && bn_sums && dz_a && dz_b, "lfd_backward: null pointer argument");
  FDDM_CHECK_ARG(phase == 0 || phase == 1, "lfd_backward: phase must be 0 or 1");
  if (D % 8 != 0) {
    set_error("lfd_backward: D=%lld must be a multiple of 8 for the tensor-core contraction", (long long)D);
    return FDDM_EUNSUPPORTED;
  }
  const LfdWorkspace lay(B, T, D);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* tables = reinterpret_cast<float*>(ws + lay.off_tables);
  float* dza = reinterpret_cast<float*>(ws + lay.off_dza);
  float* dzb = reinterpret_cast<float*>(ws + lay.off_dzb);
  const int64_t TD = T * D, rows = B * T;
  if (phase == 0) {
    // tables are recomputed here: another lfd call may have reused the workspace since the forward
    if (int rc = launch_tables(sums, TD, n_batch_global, eps, tables, stream)) return rc;
    const int terms = 2;          // G is fp32: keep hi + residual for every input dtype
    const float alpha = static_cast<float>(1.0 / n_rows_global);
    UmmaOperand Za{}, Zb{}, Gk{}, Gmn{};
    Za.ptr = z_a; Za.dtype = dtype; Za.ld = D; Za.nrows = rows; Za.ncols = D; Za.mn_is_col = 0;
    Za.scale = tables; Za.shift = tables + TD; Za.T = static_cast<int>(T); Za.stat_ld = D;
    Zb = Za;
    Zb.ptr = z_b; Zb.scale = tables + 2 * TD; Zb.shift = tables + 3 * TD;
    Gk.ptr = G; Gk.dtype = FDDM_F32; Gk.ld = D; Gk.nrows = D; Gk.ncols = D; Gk.mn_is_col = 0;   // B(n=j, k) = G[j][k]
    Gk.scale = nullptr; Gk.shift = nullptr; Gk.T = 1; Gk.stat_ld = 0;
    Gmn = Gk; Gmn.mn_is_col = 1;                                                                // B(n=k, j) = G[j][k]
    // dza~[r][j] = (1/N) sum_k zb~[r][k] G[j][k]          (oracle: B2 @ G.T / N)
    if (int rc = umma_gemm(Zb, Gk, rows, D, D, 1, terms, alpha, dza, D, 0, stream)) return rc;
    // dzb~[r][k] = (1/N) sum_j za~[r][j] G[j][k]          (oracle: A2 @ G / N)
    if (int rc = umma_gemm(Za, Gmn, rows, D, D, 1, terms, alpha, dzb, D, 0, stream)) return rc;
    return FDDM_DISPATCH_DT(dtype, launch_bn, z_a, z_b, dtype, B, TD, dza, dzb, tables, bn_sums, stream);
  }
  return FDDM_DISPATCH_DT(dtype, launch_finalize, z_a, z_b, dtype, B, TD, dza, dzb, tables, bn_sums,
                          1.0 / n_batch_global, grad_scale, dz_a, dz_b, stream);
}

}  // extern "C"
