// common.cuh -- device/host helpers shared by the sm_100a kernels of libfddm_b200.
//
// Row kernels in this library follow one pattern: a persistent CTA walks token rows; each row of V
// vocab entries is fetched once from HBM by a 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) into a
// shared-memory ring guarded by mbarriers, converted to fp32 registers with 128-bit LDS, and all
// passes of the row math (max, exp-sum, posterior/KL terms, argmax, gradient) run out of registers
// with warp-shuffle + shared-memory block reductions.  Rows whose byte length is not a multiple of
// 16 (or that are too long for the register file) take the generic path: the row is staged as fp32
// in shared memory by plain loads and the same row math runs on that copy.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fddm_b200.h"

namespace fddm {

// ------------------------------------------------------------------------------------------------
// host side: error reporting + launch accounting
// ------------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();
int row_kernel_sms();   // num_sms() minus the SMs reserved for a concurrent collective (fddm_set_sm_reserve)

#define FDDM_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      ::fddm::set_error(__VA_ARGS__);        \
      return FDDM_EINVAL;                    \
    }                                        \
  } while (0)

#define FDDM_CUDA_OK(expr)                                                             \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::fddm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FDDM_ECUDA;                                                               \
    }                                                                                  \
  } while (0)

#define FDDM_LAUNCH_OK()                                                               \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      ::fddm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FDDM_ECUDA;                                                               \
    }                                                                                  \
    ::fddm::count_launch();                                                            \
  } while (0)

// NVTX range around every extern "C" entry point (visible in nsys / ncu timelines; a no-op costing one
// predictable branch when no tool is attached).
struct ApiRange {
  explicit ApiRange(const char* name);
  ~ApiRange();
};
#define FDDM_API_RANGE() ::fddm::ApiRange _fddm_api_range_(__func__)

// Declared right before a kernel launch: an NVTX range named after the kernel and, while
// fddm_profile_enable(1) is in effect (bench.py's per-kernel roofline pass), a pair of CUDA events on the
// launching stream bracketing exactly that launch.  Nothing is recorded while the stream is being captured.
struct KernelScope {
  KernelScope(const char* name, cudaStream_t stream);
  ~KernelScope();
  const char* name_;
  cudaStream_t stream_;
  cudaEvent_t e0_;
};

static inline size_t dtype_size(int dtype) { return dtype == FDDM_F32 ? 4 : 2; }
static inline bool dtype_valid(int dtype) { return dtype == FDDM_F32 || dtype == FDDM_BF16 || dtype == FDDM_F16; }

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// PTX wrappers: shared-memory addresses, mbarrier, 1-D TMA bulk copy
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

// tx-count only (no arrival): lets the producer issue its copies first and arrive later, once the
// stage's metadata has been written
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// producer-side wait: a single lane spinning at full rate steals issue slots from the consumers
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(256);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier among the first `NT` "consumer" threads of a CTA (the producer warp never joins)
template <int NT>
__device__ __forceinline__ void consumer_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// global -> shared bulk copy completing on an mbarrier.  dst, src 16-byte aligned; bytes % 16 == 0.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// streaming 128-bit global accesses (data touched exactly once: keep it out of L1)
__device__ __forceinline__ uint4 ldg_stream_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// MUFU approximations (2 ulp-class): exp2 and reciprocal
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------
// dtype traits: 16-byte vectors <-> fp32
// ------------------------------------------------------------------------------------------------
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
  __device__ static __forceinline__ float round_trip(float x) { return x; }
  __device__ static __forceinline__ float load1(const float* p) { return *p; }
  __device__ static __forceinline__ void store1(float* p, float x) { *p = x; }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ static __forceinline__ float round_trip(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
  __device__ static __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  __device__ static __forceinline__ void store1(__nv_bfloat16* p, float x) { *p = __float2bfloat16_rn(x); }
};
template <>
struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
  __device__ static __forceinline__ float round_trip(float x) { return __half2float(__float2half_rn(x)); }
  __device__ static __forceinline__ float load1(const __half* p) { return __half2float(*p); }
  __device__ static __forceinline__ void store1(__half* p, float x) { *p = __float2half_rn(x); }
};

// ------------------------------------------------------------------------------------------------
// block reductions: warp shuffle, one shared-memory hop across warps, ONE named barrier each.
// The scratch is a ring of three regions used round-robin (all threads call in lockstep): a region is
// rewritten only two reductions -- hence two barriers -- after it was read, so no trailing barrier
// is needed to protect it.  Every thread gets the result.  Deterministic (fixed tree).
// ------------------------------------------------------------------------------------------------
constexpr int kRedRegion = 128;                 // floats per region: up to 4 values x 32 warps
constexpr int kRedFloats = 3 * kRedRegion;
struct RedRing {
  float* base;
  int idx;
  __device__ __forceinline__ float* next() {
    float* r = base + idx * kRedRegion;
    idx = (idx == 2) ? 0 : idx + 1;
    return r;
  }
};

template <int NT>
__device__ __forceinline__ float block_max(float v, RedRing& rr) {
  constexpr int NW = NT / 32;
  float* red = rr.next();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  consumer_sync<NT>();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) r = fmaxf(r, red[w]);
  return r;
}

template <int NT, int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], RedRing& rr) {
  constexpr int NW = NT / 32;
  static_assert(NV <= 4, "region holds 4 values per warp");
  float* red = rr.next();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[i * 32 + (threadIdx.x >> 5)] = v[i];
  }
  consumer_sync<NT>();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float r = red[i * 32];
#pragma unroll
    for (int w = 1; w < NW; ++w) r += red[i * 32 + w];
    v[i] = r;
  }
}

// Softmax statistics in one reduction: in (m, s) = this thread's max and sum_k exp(x_k - m);
// out (m, s) = the row max and sum_k exp(x_k - row max).  Pairs combine as
// (m1,s1) + (m2,s2) = (max, s1*2^((m1-max)*log2e) + s2*2^((m2-max)*log2e)).
template <int NT>
__device__ __forceinline__ void block_softmax_stats(float& m, float& s, RedRing& rr) {
  constexpr int NW = NT / 32;
  constexpr float kLog2e = 1.4426950408889634f;
  float* red = rr.next();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o), os = __shfl_xor_sync(0xffffffffu, s, o);
    const float nm = fmaxf(m, om);
    s = s * ex2_approx((m - nm) * kLog2e) + os * ex2_approx((om - nm) * kLog2e);
    m = nm;
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = m; red[32 + (threadIdx.x >> 5)] = s; }
  consumer_sync<NT>();
  float gm = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) gm = fmaxf(gm, red[w]);
  float gs = 0.0f;
#pragma unroll
  for (int w = 0; w < NW; ++w) gs = fmaf(red[32 + w], ex2_approx((red[w] - gm) * kLog2e), gs);
  m = gm; s = gs;
}

// argmax with torch semantics: larger value wins, equal values -> lower index.
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

template <int NT>
__device__ __forceinline__ void block_argmax(float& v, int& idx, RedRing& rr) {
  constexpr int NW = NT / 32;
  float* red = rr.next();
  int* redi = reinterpret_cast<int*>(red) + 32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    argmax_combine(v, idx, ov, oi);
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = v; redi[threadIdx.x >> 5] = idx; }
  consumer_sync<NT>();
  float r = red[0]; int ri = redi[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) argmax_combine(r, ri, red[w], redi[w]);
  v = r; idx = ri;
}

// one value from one thread to all (1 barrier)
template <int NT>
__device__ __forceinline__ int block_broadcast_int(bool is_source, int value, RedRing& rr) {
  int* red = reinterpret_cast<int*>(rr.next());
  if (is_source) red[0] = value;
  consumer_sync<NT>();
  return red[0];
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011) for in-kernel noise; counter = (lo, hi, offset_lo, offset_hi)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
// uint32 -> Exp(1) variate: u in (0,1], E = -log(u)
__device__ __forceinline__ float exp1_from_bits(uint32_t x) {
  float u = (static_cast<float>(x >> 8) + 1.0f) * (1.0f / 16777216.0f);
  return -__logf(u);
}
#endif  // __CUDACC__

}  // namespace fddm
