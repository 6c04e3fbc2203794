// dense_kernels.cu -- the scheduler's dense [B,L,K] probability API and the multi-step coefficients.
//
//   q_sample                 reference sched:31-50      8 B / element  (read 4 + write 4)
//   q_posterior              reference sched:52-104    12 B / element  (read 2x4 + write 4)
//   q_posterior_multi_step   reference sched:106-208   12 B / element, coefficients by a tiny kernel
//
// These keep the reference's signatures (dense fp32 probability tensors in, dense fp32 out); the hot
// training / sampling paths never call them (they work on ids: see sample_kernels.cu, kl_kernels.cu).
// One CTA per row; rows up to 8192 entries are register-resident (inputs read once from HBM), longer
// rows are re-read through L2 for the later passes.  Elementwise steps use the reference's op order
// with non-contracted fp32 intrinsics.
#include <algorithm>

#include "common.cuh"

namespace fddm {
namespace {

constexpr int kNT = 256;
constexpr int kEPT = 32;     // register-resident up to kNT*kEPT = 8192 entries per row

// Row accessor: REG=true keeps the row in kEPT registers per thread (read once from HBM); REG=false
// re-reads through L2.  Thread `tid` owns entries tid, tid+NT, ... (coalesced).
template <bool REG>
struct RowIO {
  const float* src;
  float r[REG ? kEPT : 1];
  __device__ __forceinline__ void init(const float* s, int K) {
    src = s;
    if (REG) {
#pragma unroll
      for (int i = 0; i < kEPT; ++i) {
        const int k = threadIdx.x + i * kNT;
        r[i] = (k < K) ? __ldg(src + k) : 0.0f;
      }
    }
  }
  __device__ __forceinline__ float get(int k, int i) const { return REG ? r[REG ? i : 0] : __ldg(src + k); }
};

// iterate the row: BODY sees k (entry index) and i (register slot, compile-time on the REG path)
#define FDDM_ROW_LOOP(K_, ...)                                                   \
  if (REG) {                                                                     \
    _Pragma("unroll") for (int i = 0; i < kEPT; ++i) {                           \
      const int k = threadIdx.x + i * kNT;                                       \
      if (k < (K_)) { __VA_ARGS__ }                                                   \
    }                                                                            \
  } else {                                                                       \
    for (int k = threadIdx.x; k < (K_); k += kNT) {                              \
      constexpr int i = 0;                                                       \
      __VA_ARGS__                                                                \
    }                                                                            \
  }

__device__ __forceinline__ long long clamp_t(long long t, int T) { return t < 1 ? 1 : (t > T ? T : t); }

template <bool REG>
__global__ void __launch_bounds__(kNT) q_sample_dense_kernel(const float* __restrict__ x0, const int64_t* __restrict__ t,
                                                             const float* __restrict__ alpha_bar, int T, int L, int K,
                                                             int rows, float eps, float u, float* __restrict__ out) {
  __shared__ float s_red[kRedFloats];
  RedRing red{s_red, 0};
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const float ab = alpha_bar[clamp_t(t[row / L], T) - 1];
    const float om = __fmul_rn(__fsub_rn(1.0f, ab), u);                       // (1-abar)*u      sched:47
    const float* src = x0 + static_cast<size_t>(row) * K;
    float* dst = out + static_cast<size_t>(row) * K;
    RowIO<REG> io;
    io.init(src, K);
    float s[1] = {0.0f};
    FDDM_ROW_LOOP(K, {
      const float v = fmaxf(__fadd_rn(__fmul_rn(ab, io.get(k, i)), om), eps);   // sched:47-48
      s[0] += v;
    })
    block_sum<kNT, 1>(s, red);
    const float den = fmaxf(s[0], eps);
    FDDM_ROW_LOOP(K, {
      const float v = fmaxf(__fadd_rn(__fmul_rn(ab, io.get(k, i)), om), eps);
      dst[k] = __fdiv_rn(v, den);                                             // sched:49
    })
  }
}

template <bool MULTI, bool REG>
__global__ void __launch_bounds__(kNT) q_posterior_dense_kernel(const float* __restrict__ xt, const float* __restrict__ xh,
                                                                const int64_t* __restrict__ t,
                                                                const float* __restrict__ betas,
                                                                const float* __restrict__ coeffs, int T, int B, int L,
                                                                int K, int rows, float eps, float* __restrict__ out) {
  __shared__ float s_red[kRedFloats];
  RedRing red{s_red, 0};
  const float Kf = static_cast<float>(K);
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int b = row / L;
    const float* pxt = xt + static_cast<size_t>(row) * K;
    const float* pxh = xh + static_cast<size_t>(row) * K;
    float* dst = out + static_cast<size_t>(row) * K;
    float a1, b1, a2, b2;
    bool identity = false;
    if (MULTI) {
      a1 = coeffs[b]; b1 = coeffs[B + b]; a2 = coeffs[2 * B + b]; b2 = coeffs[3 * B + b];
      identity = coeffs[4 * B] != 0.0f;
    } else {
      const long long tt = clamp_t(t[b], T);
      const float beta_t = betas[tt - 1];
      const float beta_p = (tt == 1) ? 0.0f : betas[tt - 2];                  // sched:76-86
      a1 = __fsub_rn(1.0f, beta_t); b1 = __fdiv_rn(beta_t, Kf);
      a2 = __fsub_rn(1.0f, beta_p); b2 = __fdiv_rn(beta_p, Kf);
    }
    if (identity) {                                                           // sched:133-134
      for (int k = threadIdx.x; k < K; k += kNT) dst[k] = pxt[k];
      continue;
    }
    RowIO<REG> ixt, ixh;
    ixt.init(pxt, K);
    ixh.init(pxh, K);
    float s3[3] = {0.0f, 0.0f, 0.0f};                                         // sum xt, sum xh, dot
    FDDM_ROW_LOOP(K, {
      const float x = ixt.get(k, i), h = ixh.get(k, i);
      s3[0] += x; s3[1] += h; s3[2] += __fmul_rn(x, h);
    })
    block_sum<kNT, 3>(s3, red);
    float cA, cB, denom;
    if (MULTI) {
      cA = __fmul_rn(b1, s3[0]);                                              // b_cum*sum_xt      sched:187
      cB = __fmul_rn(b2, s3[1]);                                              // b_tgt*sum_x0hat   sched:191
      denom = __fadd_rn(__fmul_rn(a1, s3[2]), __fmul_rn(__fmul_rn(b1, s3[1]), s3[0]));   // sched:197
    } else {
      cA = b1; cB = b2;                                                       // sched:91,94
      denom = __fadd_rn(__fmul_rn(a1, s3[2]), b1);                            // sched:98
    }
    denom = fmaxf(denom, eps);
    float s[1] = {0.0f};
    FDDM_ROW_LOOP(K, {
      const float A = __fadd_rn(__fmul_rn(a1, ixt.get(k, i)), cA);
      const float Bv = __fadd_rn(__fmul_rn(a2, ixh.get(k, i)), cB);
      s[0] += __fdiv_rn(__fmul_rn(A, Bv), denom);                             // sched:101 / 200
    })
    block_sum<kNT, 1>(s, red);
    const float nrm = fmaxf(s[0], eps);
    FDDM_ROW_LOOP(K, {
      const float A = __fadd_rn(__fmul_rn(a1, ixt.get(k, i)), cA);
      const float Bv = __fadd_rn(__fmul_rn(a2, ixh.get(k, i)), cB);
      dst[k] = __fdiv_rn(__fdiv_rn(__fmul_rn(A, Bv), denom), nrm);            // sched:102 / 204
    })
  }
}

// per-sample transition-product coefficients, no host sync (reference sched:132-183)
__global__ void __launch_bounds__(256) multistep_coeffs_kernel(const int64_t* __restrict__ t, int B, long long delta,
                                                               const int64_t* __restrict__ delta_cap,
                                                               const float* __restrict__ betas, int T, int K,
                                                               float* __restrict__ coeffs) {
  __shared__ long long s_min[8];
  __shared__ long long s_delta;
  long long mn = 0x7fffffffffffffffll;
  for (int b = threadIdx.x; b < B; b += blockDim.x) mn = min(mn, static_cast<long long>(t[b]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = mn;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long m = s_min[0];
    for (int w = 1; w < 8; ++w) m = min(m, s_min[w]);
    long long d = min(delta, m);                                              // sched:132 (quirk Q10)
    if (delta_cap) d = min(d, static_cast<long long>(*delta_cap));
    s_delta = d;
    coeffs[4 * B] = (d <= 0) ? 1.0f : 0.0f;                                   // sched:133-134
  }
  __syncthreads();
  const long long d = s_delta;
  const float Kf = static_cast<float>(K);
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float a = 1.0f, bb = 0.0f, a_g = 1.0f, b_g = 0.0f;
    if (d > 0) {
      const long long tc = t[b];
      long long tg = tc - d;
      if (tg < 0) tg = 0;                                                     // sched:137
      for (long long step = tc; step > tg; --step) {
        if (step >= 1 && step <= T) {
          const float beta_s = betas[step - 1];
          const float a_s = __fsub_rn(1.0f, beta_s);
          const float b_s = __fdiv_rn(beta_s, Kf);
          // quirk Q1: the reference's a_old is an aliased view, so the UPDATED a enters b's update
          a = __fmul_rn(a_s, a);                                              // sched:160
          bb = __fadd_rn(__fmul_rn(a_s, bb), __fmul_rn(b_s, __fadd_rn(a, __fmul_rn(Kf, bb))));   // sched:161
        }
      }
      if (tg > 0 && tg <= T) {                                                // sched:173-179
        a_g = __fsub_rn(1.0f, betas[tg - 1]);
        b_g = __fdiv_rn(betas[tg - 1], Kf);
      }
    }
    coeffs[b] = a; coeffs[B + b] = bb; coeffs[2 * B + b] = a_g; coeffs[3 * B + b] = b_g;
  }
}

}  // namespace
}  // namespace fddm

extern "C" {

int fddm_q_sample_dense(const float* x0_prob, const int64_t* t, const float* alpha_bar, int64_t T, int64_t B, int64_t L,
                        int64_t K, float eps, float* out, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(x0_prob && t && alpha_bar && out, "q_sample_dense: null pointer argument");
  FDDM_CHECK_ARG(B > 0 && L > 0 && K > 0 && T > 0, "q_sample_dense: non-positive size");
  FDDM_CHECK_ARG(B * L < (1ll << 31) && K < (1ll << 30), "q_sample_dense: size too large");
  const int rows = static_cast<int>(B * L);
  const int grid = static_cast<int>(std::min<int64_t>(rows, static_cast<int64_t>(num_sms()) * 8));
  const float u = static_cast<float>(1.0 / static_cast<double>(K));
  KernelScope ks("q_sample_dense_kernel", stream);
  if (K <= kNT * kEPT)
    q_sample_dense_kernel<true><<<grid, kNT, 0, stream>>>(x0_prob, t, alpha_bar, static_cast<int>(T),
                                                          static_cast<int>(L), static_cast<int>(K), rows, eps, u, out);
  else
    q_sample_dense_kernel<false><<<grid, kNT, 0, stream>>>(x0_prob, t, alpha_bar, static_cast<int>(T),
                                                           static_cast<int>(L), static_cast<int>(K), rows, eps, u, out);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int fddm_q_posterior_dense(const float* xt_prob, const float* x0hat_prob, const int64_t* t, const float* betas,
                           int64_t T, int64_t B, int64_t L, int64_t K, float eps, float* out, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(xt_prob && x0hat_prob && t && betas && out, "q_posterior_dense: null pointer argument");
  FDDM_CHECK_ARG(B > 0 && L > 0 && K > 0 && T > 0, "q_posterior_dense: non-positive size");
  FDDM_CHECK_ARG(B * L < (1ll << 31) && K < (1ll << 30), "q_posterior_dense: size too large");
  const int rows = static_cast<int>(B * L);
  const int grid = static_cast<int>(std::min<int64_t>(rows, static_cast<int64_t>(num_sms()) * 4));
  KernelScope ks("q_posterior_dense_kernel", stream);
  if (K <= kNT * kEPT)
    q_posterior_dense_kernel<false, true><<<grid, kNT, 0, stream>>>(xt_prob, x0hat_prob, t, betas, nullptr,
                                                                    static_cast<int>(T), static_cast<int>(B),
                                                                    static_cast<int>(L), static_cast<int>(K), rows, eps,
                                                                    out);
  else
    q_posterior_dense_kernel<false, false><<<grid, kNT, 0, stream>>>(xt_prob, x0hat_prob, t, betas, nullptr,
                                                                     static_cast<int>(T), static_cast<int>(B),
                                                                     static_cast<int>(L), static_cast<int>(K), rows, eps,
                                                                     out);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int fddm_multistep_coeffs(const int64_t* t, int64_t B, int64_t delta, const int64_t* delta_cap, const float* betas,
                          int64_t T, int64_t K, float* coeffs, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(t && betas && coeffs, "multistep_coeffs: null pointer argument");
  FDDM_CHECK_ARG(B > 0 && K > 0 && T > 0 && B < (1ll << 30), "multistep_coeffs: bad size");
  KernelScope ks("multistep_coeffs_kernel", stream);
  multistep_coeffs_kernel<<<1, 256, 0, stream>>>(t, static_cast<int>(B), static_cast<long long>(delta), delta_cap, betas,
                                                 static_cast<int>(T), static_cast<int>(K), coeffs);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int fddm_q_posterior_multi_dense(const float* xt_prob, const float* x0hat_prob, const float* coeffs, int64_t B,
                                 int64_t L, int64_t K, float eps, float* out, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(xt_prob && x0hat_prob && coeffs && out, "q_posterior_multi_dense: null pointer argument");
  FDDM_CHECK_ARG(B > 0 && L > 0 && K > 0, "q_posterior_multi_dense: non-positive size");
  FDDM_CHECK_ARG(B * L < (1ll << 31) && K < (1ll << 30), "q_posterior_multi_dense: size too large");
  const int rows = static_cast<int>(B * L);
  const int grid = static_cast<int>(std::min<int64_t>(rows, static_cast<int64_t>(num_sms()) * 4));
  KernelScope ks("q_posterior_multi_dense_kernel", stream);
  if (K <= kNT * kEPT)
    q_posterior_dense_kernel<true, true><<<grid, kNT, 0, stream>>>(xt_prob, x0hat_prob, nullptr, nullptr, coeffs, 0,
                                                                   static_cast<int>(B), static_cast<int>(L),
                                                                   static_cast<int>(K), rows, eps, out);
  else
    q_posterior_dense_kernel<true, false><<<grid, kNT, 0, stream>>>(xt_prob, x0hat_prob, nullptr, nullptr, coeffs, 0,
                                                                    static_cast<int>(B), static_cast<int>(L),
                                                                    static_cast<int>(K), rows, eps, out);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

}  // extern "C"
