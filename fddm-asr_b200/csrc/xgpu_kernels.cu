// xgpu_kernels.cu -- the library's own SUM all-reduce for the three small L_fd exchange buffers.
//
// The batch-sharded L_fd has three exchange steps per training step (statistics fp64 [4][T*D], covariance fp32
// [D*D], batch-norm moment fp32 [2][T*D]): 1.6 - 6.3 MB each, latency-bound for a ring/tree collective
// (NCCL: 31 - 58 us each on 8 B200s, profiles/r01b_nccl_probe_n8.txt).  With the buffers in SYMMETRIC memory
// (every rank has every peer's copy mapped over NVLink/NVSwitch) the all-reduce is one small kernel, in place:
//   algo 1, "p2p two-shot":  each rank loads ITS 1/world slice from every peer's copy (plain peer loads, summed
//           in rank order -> the result is bitwise the same on every rank), then stores the sums into every
//           peer's copy.
//   algo 2, "NVLS":          the same slice is reduced by the switch (multimem.ld_reduce.add) and re-broadcast
//           by the switch (multimem.st) through the multicast mapping.
// Both run between two cross-GPU barriers: one-way flags (st.release.sys into the peer's pad, ld.acquire.sys spin
// on the local pad) carrying a per-launch epoch that lives in device memory, so the kernel is replay-safe in a
// CUDA graph.  No copy engine, no proxy thread, no NCCL kernel.
//
// The caller owns the symmetric allocation, the peer / multicast mappings and the signal pads (plain pointers in
// this C ABI; the Python mirror gets them from torch.distributed._symmetric_memory).  The library uses the pad
// bytes [kPadOffset, kPadOffset + pad_words * 4) only, so the owner's own barriers (which start at offset 0) are
// not disturbed.
#include <algorithm>

#include "common.cuh"

namespace fddm {
namespace {

constexpr int kXThreads = 512;
constexpr int kXMaxBlocks = 32;
constexpr int kXMaxWorld = 16;
constexpr int kPadOffset = 4096;                                   // bytes; the owner's barriers use the front
// pad words (32-bit) from kPadOffset: flags [block][peer], then one epoch word per block
constexpr int kPadWords = kXMaxBlocks * kXMaxWorld + kXMaxBlocks;

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
  asm volatile("st.global.release.sys.b32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.global.acquire.sys.b32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

struct XParams {
  void* const* bufs;           // device array [world]: every rank's copy of the buffer (peer mappings)
  void* mc;                    // multicast address of the buffer (algo 2) or nullptr
  uint32_t* const* pads;       // device array [world] of signal-pad base pointers
  int rank, world;
  int64_t n;                   // elements
};

// CTA `blockIdx.x` of this rank meets CTA `blockIdx.x` of every peer.  `epoch` is this CTA's barrier count (the
// same on every rank because every rank makes the same calls): thread p (< world) writes it into peer p's pad and
// waits until peer p's write has arrived in the local pad.  Signed distance: correct across the 32-bit wrap.
__device__ __forceinline__ void xgpu_barrier(const XParams& p, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < p.world) {
    const int peer = threadIdx.x;
    __threadfence_system();                                        // this CTA's earlier stores, system-wide
    uint32_t* put = p.pads[peer] + kPadOffset / 4 + blockIdx.x * kXMaxWorld + p.rank;
    st_release_sys(put, epoch);
    const uint32_t* take = p.pads[p.rank] + kPadOffset / 4 + blockIdx.x * kXMaxWorld + peer;
    while (static_cast<int32_t>(ld_acquire_sys(take) - epoch) < 0) {}
  }
  __syncthreads();
}

// this CTA's epoch counter: read, and advanced by 2 (two barriers per launch) by thread 0 after the last barrier
__device__ __forceinline__ uint32_t xgpu_epoch_begin(const XParams& p) {
  return *(p.pads[p.rank] + kPadOffset / 4 + kXMaxBlocks * kXMaxWorld + blockIdx.x);
}
__device__ __forceinline__ void xgpu_epoch_end(const XParams& p, uint32_t e0) {
  if (threadIdx.x == 0) *(p.pads[p.rank] + kPadOffset / 4 + kXMaxBlocks * kXMaxWorld + blockIdx.x) = e0 + 2;
}

template <typename V>
struct Unit;                                                       // 16-byte unit of the element type
template <>
struct Unit<float> {
  using type = float4;
  static constexpr int N = 4;
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ void add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
};
template <>
struct Unit<double> {
  using type = double2;
  static constexpr int N = 2;
  static __device__ __forceinline__ double2 zero() { return make_double2(0.0, 0.0); }
  static __device__ __forceinline__ void add(double2& a, const double2& b) { a.x += b.x; a.y += b.y; }
};

// algo 1: peer loads of this rank's slice from every copy (summed in rank order), peer stores of the sums.
// A thread keeps WORLD * U 16-byte peer loads in flight (32 / WORLD units each: the same 512 bytes per thread for
// every world size -- the slice of a small world is longer and needs the deeper queue to cover the NVLink round trip).
template <typename V, int WORLD>
__global__ void __launch_bounds__(kXThreads) xgpu_allreduce_p2p_kernel(const XParams p) {
  using U16 = typename Unit<V>::type;
  constexpr int U = 32 / WORLD / 4 > 0 ? 32 / WORLD / 4 : 1;       // world 2: 4, world 4: 2, world 8: 1
  const uint32_t e0 = xgpu_epoch_begin(p);
  xgpu_barrier(p, e0 + 1);
  const int64_t units = p.n / Unit<V>::N;
  const int64_t per = (units + WORLD - 1) / WORLD;
  const int64_t u0 = per * p.rank, u1 = min(units, u0 + per);
  U16* buf[WORLD];
#pragma unroll
  for (int r = 0; r < WORLD; ++r) buf[r] = static_cast<U16*>(p.bufs[r]);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kXThreads;
  for (int64_t i = u0 + static_cast<int64_t>(blockIdx.x) * kXThreads + threadIdx.x; i < u1; i += U * stride) {
    U16 v[U][WORLD];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int64_t j = i + k * stride;
#pragma unroll
      for (int r = 0; r < WORLD; ++r)
        if (j < u1) v[k][r] = __ldcg(buf[r] + j);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int64_t j = i + k * stride;
      if (j < u1) {
        U16 acc = v[k][0];
#pragma unroll
        for (int r = 1; r < WORLD; ++r) Unit<V>::add(acc, v[k][r]);
#pragma unroll
        for (int r = 0; r < WORLD; ++r) __stcg(buf[r] + j, acc);
      }
    }
  }
  xgpu_barrier(p, e0 + 2);
  xgpu_epoch_end(p, e0);
}

// algo 2: the switch reduces (multimem.ld_reduce) and broadcasts (multimem.st)
template <typename V>
__global__ void __launch_bounds__(kXThreads) xgpu_allreduce_nvls_kernel(const XParams p) {
  const uint32_t e0 = xgpu_epoch_begin(p);
  xgpu_barrier(p, e0 + 1);
  constexpr int N = (sizeof(V) == 4) ? 4 : 1;                      // fp32: v4 forms; fp64: scalar forms only
  const int64_t units = p.n / N;
  const int64_t per = (units + p.world - 1) / p.world;
  const int64_t u0 = per * p.rank, u1 = min(units, u0 + per);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kXThreads;
  constexpr int U = (sizeof(V) == 4) ? 4 : 8;                      // independent reductions in flight per thread
  for (int64_t i = u0 + static_cast<int64_t>(blockIdx.x) * kXThreads + threadIdx.x; i < u1; i += U * stride) {
    if constexpr (sizeof(V) == 4) {
      float4* base = static_cast<float4*>(p.mc);
      float4 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int64_t j = i + k * stride;
        if (j < u1)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w)
                       : "l"(base + j)
                       : "memory");
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int64_t j = i + k * stride;
        if (j < u1)
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(base + j), "f"(v[k].x),
                       "f"(v[k].y), "f"(v[k].z), "f"(v[k].w)
                       : "memory");
      }
    } else {
      double* base = static_cast<double*>(p.mc);
      double v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int64_t j = i + k * stride;
        if (j < u1)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v[k]) : "l"(base + j) : "memory");
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int64_t j = i + k * stride;
        if (j < u1) asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(base + j), "d"(v[k]) : "memory");
      }
    }
  }
  xgpu_barrier(p, e0 + 2);
  xgpu_epoch_end(p, e0);
}

template <typename V>
int launch_p2p(const XParams& p, int blocks, cudaStream_t stream) {
  switch (p.world) {
    case 2: xgpu_allreduce_p2p_kernel<V, 2><<<blocks, kXThreads, 0, stream>>>(p); break;
    case 4: xgpu_allreduce_p2p_kernel<V, 4><<<blocks, kXThreads, 0, stream>>>(p); break;
    case 8: xgpu_allreduce_p2p_kernel<V, 8><<<blocks, kXThreads, 0, stream>>>(p); break;
    default:
      set_error("xgpu_allreduce: the p2p algorithm is built for world 2, 4 and 8 (got %d)", p.world);
      return FDDM_EUNSUPPORTED;
  }
  return FDDM_OK;
}

}  // namespace
}  // namespace fddm

extern "C" {

int64_t fddm_xgpu_signal_pad_bytes(void) {
  return static_cast<int64_t>(fddm::kPadOffset) + static_cast<int64_t>(fddm::kPadWords) * 4;
}

int fddm_xgpu_allreduce(const void* buffer_ptrs_dev, void* multicast_ptr, const void* signal_pad_ptrs_dev, int rank,
                        int world, int elem_bytes, int64_t n, int algo, int max_ctas, fddm_stream_t stream_) {
  using namespace fddm;
  FDDM_API_RANGE();
  FDDM_CHECK_ARG(signal_pad_ptrs_dev != nullptr, "xgpu_allreduce: null signal pads");
  FDDM_CHECK_ARG(world >= 2 && world <= kXMaxWorld && rank >= 0 && rank < world, "xgpu_allreduce: rank %d of world %d",
                 rank, world);
  FDDM_CHECK_ARG(elem_bytes == 4 || elem_bytes == 8, "xgpu_allreduce: elem_bytes must be 4 (fp32) or 8 (fp64)");
  FDDM_CHECK_ARG(n > 0 && n % (16 / elem_bytes) == 0, "xgpu_allreduce: the length must be a positive multiple of 16 bytes");
  FDDM_CHECK_ARG(algo == FDDM_XGPU_P2P || algo == FDDM_XGPU_NVLS, "xgpu_allreduce: unknown algorithm %d", algo);
  FDDM_CHECK_ARG(max_ctas >= 0, "xgpu_allreduce: max_ctas must be >= 0");
  FDDM_CHECK_ARG(algo != FDDM_XGPU_P2P || buffer_ptrs_dev != nullptr, "xgpu_allreduce: p2p needs the peer buffer pointers");
  FDDM_CHECK_ARG(algo != FDDM_XGPU_NVLS || (multicast_ptr != nullptr && reinterpret_cast<uintptr_t>(multicast_ptr) % 16 == 0),
                 "xgpu_allreduce: NVLS needs a 16-byte aligned multicast address");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XParams p;
  p.bufs = static_cast<void* const*>(buffer_ptrs_dev);
  p.mc = multicast_ptr;
  p.pads = static_cast<uint32_t* const*>(signal_pad_ptrs_dev);
  p.rank = rank; p.world = world; p.n = n;
  // enough CTAs to keep the rank's whole slice in flight (an NVLink round trip is ~2 us), few enough to run next
  // to a persistent row kernel that left SMs free (fddm_set_sm_reserve).  Same value on every rank.
  const int64_t units = n / (16 / elem_bytes);
  const int64_t per_rank = (units + world - 1) / world;
  const int per_thread = (algo == FDDM_XGPU_P2P) ? std::max(1, 8 / world) : 4;
  const int cap = (max_ctas > 0) ? std::min(max_ctas, kXMaxBlocks) : kXMaxBlocks;
  const int blocks = static_cast<int>(std::max<int64_t>(
      1, std::min<int64_t>(cap, (per_rank + static_cast<int64_t>(kXThreads) * per_thread - 1) /
                                            (static_cast<int64_t>(kXThreads) * per_thread))));
  int rc = FDDM_OK;
  if (algo == FDDM_XGPU_P2P) {
    KernelScope ks(elem_bytes == 4 ? "xgpu_allreduce_p2p_f32" : "xgpu_allreduce_p2p_f64", stream);
    rc = (elem_bytes == 4) ? launch_p2p<float>(p, blocks, stream) : launch_p2p<double>(p, blocks, stream);
    if (rc != FDDM_OK) return rc;
    FDDM_LAUNCH_OK();
  } else {
    KernelScope ks(elem_bytes == 4 ? "xgpu_allreduce_nvls_f32" : "xgpu_allreduce_nvls_f64", stream);
    if (elem_bytes == 4) xgpu_allreduce_nvls_kernel<float><<<blocks, kXThreads, 0, stream>>>(p);
    else xgpu_allreduce_nvls_kernel<double><<<blocks, kXThreads, 0, stream>>>(p);
    FDDM_LAUNCH_OK();
  }
  return FDDM_OK;
}

}  // extern "C"
