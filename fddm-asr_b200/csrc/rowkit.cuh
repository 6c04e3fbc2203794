// rowkit.cuh -- row stores and the TMA producer/consumer ring used by every token-row kernel.
//
// CTA shape of a "ring" kernel: NT consumer threads (threadIdx.x < NT) + one producer warp
// (threadIdx.x >= NT).  The producer warp claims token rows from a global work counter (dynamic
// scheduling: masked rows cost nothing, so static striding would leave SMs idle), prepares the
// row's metadata, and issues one cp.async.bulk per input stream into the next free ring stage.
// Consumers wait on the stage's `full` mbarrier, pull the row into registers, hand the stage back
// through the `empty` mbarrier (one arrive per consumer warp) and do all the math from registers
// while the producer is already fetching rows for later iterations.
#pragma once

#include "common.cuh"

namespace fddm {

constexpr int kMaxStages = 4;
constexpr float kNegInf = -3.0e38f;   // finite stand-in for -inf in padded lanes (no inf-inf NaNs)

// ------------------------------------------------------------------------------------------------
// RegRow: the row lives in EPT fp32 registers per thread.  Thread `tid` owns the 16-byte vectors
// tid, tid+NT, tid+2NT, ... of the row; lanes past the end of the row are inactive.
// ------------------------------------------------------------------------------------------------
// TIGHT: the launcher guarantees nvec > (NVEC-1)*NT, i.e. every thread's first NVEC-1 vectors exist and only
// the last one needs a bounds check -- the per-vector predicates (ISETP/BRA/BSSY/BSYNC around every unrolled
// body, ~18 % of the KL kernel's instructions) disappear from all but one vector per pass.
template <typename T, int NT, int EPT, bool TIGHT = false>
struct RegRow {
  static constexpr int N = Vec16<T>::N;       // elements per 16-byte vector
  static constexpr int NVEC = EPT / N;
  static_assert(EPT % N == 0, "EPT must be a whole number of vectors");
  float v[EPT];
  int nvec;                                   // vectors in the row (V / N)
  int tid;
  __device__ __forceinline__ bool has(int j, int vi) const {
    if (TIGHT && j < NVEC - 1) return true;
    return vi < nvec;
  }

  __device__ __forceinline__ void load_from_smem(const void* stage, int V, int tid_) {
    tid = tid_;
    nvec = V / N;
    const uint4* s = reinterpret_cast<const uint4*>(stage);
#pragma unroll
    for (int j = 0; j < NVEC; ++j) {
      const int vi = j * NT + tid;
      if (has(j, vi)) {
        Vec16<T>::unpack(s[vi], &v[j * N]);
      } else {
#pragma unroll
        for (int e = 0; e < N; ++e) v[j * N + e] = kNegInf;
      }
    }
  }
  // f(k, x&) over the valid elements owned by this thread, ascending k within the thread
  template <class F>
  __device__ __forceinline__ void for_each(F&& f) {
#pragma unroll
    for (int j = 0; j < NVEC; ++j) {
      const int vi = j * NT + tid;
      if (has(j, vi)) {
#pragma unroll
        for (int e = 0; e < N; ++e) f(vi * N + e, v[j * N + e]);
      }
    }
  }
  // read-only visit f(k, x) (same as for_each here; rows that live in shared memory skip the write-back)
  template <class F>
  __device__ __forceinline__ void for_each_ro(F&& f) {
    for_each([&](int k, float& x) { f(k, x); });
  }
  // which thread owns element k / how many elements this thread owns
  __device__ __forceinline__ int owner_of(int k) const { return (k / N) % NT; }
  __device__ __forceinline__ int n_owned() const {
    const int nv = (nvec - tid + NT - 1) / NT;
    return (nv > 0 ? nv : 0) * N;
  }
  // f4(x4) over groups of 4 consecutive valid elements (x4 points at 4 registers)
  template <class F4, class F1>
  __device__ __forceinline__ void for_each4(F4&& f4, F1&&) {
#pragma unroll
    for (int j = 0; j < NVEC; ++j) {
      const int vi = j * NT + tid;
      if (has(j, vi)) {
#pragma unroll
        for (int q = 0; q < N / 4; ++q) f4(&v[j * N + 4 * q]);
      }
    }
  }
  // dst[k..k+3] = g4(x4, o4) for the whole row, 128-bit streaming stores
  template <class G4, class G1>
  __device__ __forceinline__ void store4(T* dst, G4&& g4, G1&&) {
#pragma unroll
    for (int j = 0; j < NVEC; ++j) {
      const int vi = j * NT + tid;
      if (has(j, vi)) {
        float o[N];
#pragma unroll
        for (int q = 0; q < N / 4; ++q) g4(&v[j * N + 4 * q], &o[4 * q]);
        stg_stream_v4(reinterpret_cast<uint4*>(dst) + vi, Vec16<T>::pack(o));
      }
    }
  }
  // same, g4(k0, x4, o4) also gets the index of the group's first element
  template <class G4, class G1>
  __device__ __forceinline__ void store4k(T* dst, G4&& g4, G1&&) {
#pragma unroll
    for (int j = 0; j < NVEC; ++j) {
      const int vi = j * NT + tid;
      if (has(j, vi)) {
        float o[N];
#pragma unroll
        for (int q = 0; q < N / 4; ++q) g4(vi * N + 4 * q, &v[j * N + 4 * q], &o[4 * q]);
        stg_stream_v4(reinterpret_cast<uint4*>(dst) + vi, Vec16<T>::pack(o));
      }
    }
  }
  // dst[k] = g(k, x) for the whole row, 128-bit streaming stores
  template <class G>
  __device__ __forceinline__ void store(T* dst, G&& g) {
#pragma unroll
    for (int j = 0; j < NVEC; ++j) {
      const int vi = j * NT + tid;
      if (has(j, vi)) {
        float o[N];
#pragma unroll
        for (int e = 0; e < N; ++e) o[e] = g(vi * N + e, v[j * N + e]);
        stg_stream_v4(reinterpret_cast<uint4*>(dst) + vi, Vec16<T>::pack(o));
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// SmemRow: generic path, the row is an fp32 array in shared memory (any V, any alignment).
// ------------------------------------------------------------------------------------------------
template <typename T, int NT>
struct SmemRow {
  float* r;
  int V;
  int tid;
  __device__ __forceinline__ void load_from_gmem(float* smem_row, const T* src, int V_, int tid_) {
    r = smem_row; V = V_; tid = tid_;
    for (int k = tid; k < V; k += NT) r[k] = Vec16<T>::load1(src + k);
    consumer_sync<NT>();
  }
  // Element ownership is the same in every visitor (kernels keep per-thread state such as a local
  // softmax max between passes): thread t owns the 4-element groups t, t+NT, ... and, of the V % 4
  // tail, element V4 + t.
  template <class F>
  __device__ __forceinline__ void for_each(F&& f) {
    const int V4 = V & ~3;
    for (int k = 4 * tid; k < V4; k += 4 * NT) {
#pragma unroll
      for (int e = 0; e < 4; ++e) f(k + e, r[k + e]);
    }
    for (int k = V4 + tid; k < V; k += NT) f(k, r[k]);
  }
  template <class F>
  __device__ __forceinline__ void for_each_ro(F&& f) {
    for_each([&](int k, float& x) { f(k, x); });
  }
  template <class G>
  __device__ __forceinline__ void store(T* dst, G&& g) {
    const int V4 = V & ~3;
    for (int k = 4 * tid; k < V4; k += 4 * NT) {
#pragma unroll
      for (int e = 0; e < 4; ++e) Vec16<T>::store1(dst + k + e, g(k + e, r[k + e]));
    }
    for (int k = V4 + tid; k < V; k += NT) Vec16<T>::store1(dst + k, g(k, r[k]));
  }
  __device__ __forceinline__ int owner_of(int k) const {
    const int V4 = V & ~3;
    return k < V4 ? (k >> 2) % NT : k - V4;
  }
  __device__ __forceinline__ int n_owned() const {
    const int V4 = V & ~3, g = V4 >> 2;
    const int ng = (g - tid + NT - 1) / NT;
    return (ng > 0 ? ng : 0) * 4 + ((V4 + tid < V) ? 1 : 0);
  }
  // groups of 4 consecutive elements, then the (V % 4) tail one by one
  template <class F4, class F1>
  __device__ __forceinline__ void for_each4(F4&& f4, F1&& f1) {
    const int V4 = V & ~3;
    for (int k = 4 * tid; k < V4; k += 4 * NT) f4(&r[k]);
    for (int k = V4 + tid; k < V; k += NT) f1(r[k]);
  }
  template <class G4, class G1>
  __device__ __forceinline__ void store4(T* dst, G4&& g4, G1&& g1) {
    const int V4 = V & ~3;
    for (int k = 4 * tid; k < V4; k += 4 * NT) {
      float o[4];
      g4(&r[k], o);
#pragma unroll
      for (int e = 0; e < 4; ++e) Vec16<T>::store1(dst + k + e, o[e]);
    }
    for (int k = V4 + tid; k < V; k += NT) Vec16<T>::store1(dst + k, g1(r[k]));
  }
  // index-aware variants: g4(k0, x4, o4), g1(k, x)
  template <class G4, class G1>
  __device__ __forceinline__ void store4k(T* dst, G4&& g4, G1&& g1) {
    const int V4 = V & ~3;
    for (int k = 4 * tid; k < V4; k += 4 * NT) {
      float o[4];
      g4(k, &r[k], o);
#pragma unroll
      for (int e = 0; e < 4; ++e) Vec16<T>::store1(dst + k + e, o[e]);
    }
    for (int k = V4 + tid; k < V; k += NT) Vec16<T>::store1(dst + k, g1(k, r[k]));
  }
};

// ------------------------------------------------------------------------------------------------
// Ring bookkeeping shared by producer and consumers
// ------------------------------------------------------------------------------------------------
struct RingMeta {          // one per stage, written by the producer before it arms `full`
  int row;                 // token row index, -1 = no more work
  float w;                 // row weight (0 => nothing was loaded for this row)
  int i0, i1;              // per-kernel integers (token ids)
  float f0, f1, f2, f3;    // per-kernel scalars (schedule coefficients)
  uint32_t r0, r1, r2;     // per-row random bits drawn by the producer (in-kernel RNG kernels)
};

struct Ring {
  uint8_t* stages;         // nstages * stage_bytes, 128-byte aligned
  uint32_t stage_bytes;
  int nstages;
  uint64_t* full;          // [kMaxStages]
  uint64_t* empty;         // [kMaxStages]
  RingMeta* meta;          // [kMaxStages]

  __device__ __forceinline__ uint8_t* stage(int s) const { return stages + static_cast<size_t>(s) * stage_bytes; }
};

// Called by all threads of the CTA (consumers + producer warp) before the role split.
template <int NT>
__device__ __forceinline__ void ring_init(Ring& ring) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < ring.nstages; ++s) {
      mbar_init(&ring.full[s], 1);
      mbar_init(&ring.empty[s], NT / 32);
    }
    mbar_fence_init();
    fence_proxy_async();
  }
  __syncthreads();
}

// Consumer side: hand a stage back to the producer once this warp has finished reading it.
__device__ __forceinline__ void ring_release(Ring& ring, int s) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&ring.empty[s]);
}

// host: choose the ring depth for a given per-stage footprint
struct RingPlan {
  int nstages;
  int ctas_per_sm;
  size_t smem_bytes;       // dynamic shared memory per CTA
};
// `want_ctas` > 0: resident CTAs per SM the kernel was compiled for (falls back towards 1 while a CTA would
// get fewer than two stages)
static inline RingPlan plan_ring(size_t stage_bytes, int nt_consumers, int want_ctas = 0, int min_stages = 2) {
  RingPlan p{};
  const size_t sb = (stage_bytes + 127) & ~size_t(127);
  const size_t budget_total = 216 * 1024;   // leave room for static smem + driver reservation
  int ctas = want_ctas > 0 ? want_ctas : ((nt_consumers <= 256) ? 2 : 1);
  size_t per = budget_total / ctas;
  int st = static_cast<int>(per / sb);
  while (st < min_stages && ctas > 1) { --ctas; per = budget_total / ctas; st = static_cast<int>(per / sb); }
  if (st > kMaxStages) st = kMaxStages;
  p.nstages = st;          // may be 0 -> caller falls back / reports unsupported
  p.ctas_per_sm = ctas;
  p.smem_bytes = sb * static_cast<size_t>(st > 0 ? st : 0) + 128;
  return p;
}

}  // namespace fddm
