// lfd_umma_bwd.cu -- the backward contractions of L_fd as ONE persistent tcgen05 kernel per gradient:
//
//   dz~[r][n]       = alpha * sum_k Z(r, k) * G(n, k)          Z: z~ planes of the other tensor, G: dloss/dC (or G^T)
//   partial[g][n]   = sum_{r in 32-row group g} dz~[r][n] * Zs(r, n)      Zs: z~ planes of the tensor dz~ belongs to
//
// (reference: autograd of losses/fddm_losses.py:48-57 through _standardize, losses:18-26.)
//
// What changed against the one-tile-per-CTA kernel it replaces (profiles/r01b_umma_bwd_ncu_full_selected.csv:
// tensor pipe 25 % of elapsed, 2.6 waves, a 131 KB epilogue and the TMEM-alloc prologue exposed around a
// 24-k-block main loop; then a separate pass over dz~ and z for the batch-norm sums):
//   * persistent CTAs (one per SM) walk the (m-tile, n-tile) list; the operand ring runs on across tiles;
//   * TWO 256-column TMEM accumulators alternate, so the epilogue of tile i (TMEM -> registers -> global) runs
//     under the main loop of tile i+1;
//   * rows are tb-major (lfd_common.cuh): the 32 rows an epilogue warp owns share one position t, so the batch
//     sums of dz~ * z~ are a 32-lane butterfly over registers the epilogue already holds -- the batch-norm
//     reduction pass over dz~ and z (lfd_bn_reduce_kernel) disappears; z~ comes from the packed planes with
//     fully coalesced 16-byte loads.
//
// Warp roles as in lfd_umma.cu: warp 0 TMA producer, warp 1 MMA issuer (one lane) + TMEM owner, warps 2..5
// epilogue (TMEM lane quarter = warp id % 4).
#include <stdlib.h>

#include <algorithm>

#include "umma_common.cuh"

namespace fddm {
namespace {
using namespace umma;

constexpr int kBwdBN = 256;
constexpr int kBwdStages = 4;
constexpr uint32_t kBwdStageBytes = 2 * (kTileA + kTileB);          // hi + lo planes of A and B: 48 KB

struct BwdParams {
  const __nv_bfloat16 *z_hi, *z_lo;      // A operand  [D_pad/8][R_pad][8]
  const __nv_bfloat16 *g_hi, *g_lo;      // B operand  [D_pad/8][D_pad][8]
  const __nv_bfloat16 *s_hi, *s_lo;      // statistics operand, same layout as z
  int64_t R_pad, D_pad;
  int D;                                 // N = K = D
  int tiles_n, num_kb, total;
  int Bn, Bp, Tn;
  float alpha;
  float* dz;                             // [B][T][D] fp32
  float* partial;                        // [T * Bp/32][D] fp32
};

// ---- thread-block-cluster helpers (CL = 2: the two CTAs of a cluster work on the same n-tile and each loads
// half of the G tile, multicast into both CTAs' shared memory -> the L2 -> SM operand traffic, which bounds this
// kernel, drops from 48 to 32 KB per k-block) ------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy delivered to the same shared-memory offset of every CTA in `mask`; each destination CTA's mbarrier
// at the same offset receives the complete_tx for the bytes written there
__device__ __forceinline__ void tma_load_1d_multicast(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                      uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// MMA-completion arrive on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <int CL>
__global__ void __launch_bounds__(kThreads, 1) umma_bwd_kernel(const BwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t s_full[kBwdStages], s_empty[kBwdStages], s_acc_full[2], s_acc_empty[2];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1u);

  if (tid == 0) {
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], CL);                  // one MMA-completion arrive from every CTA of the cluster
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_acc_full[b], 1);
      mbar_init(&s_acc_empty[b], kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 1) {                                   // the MMA warp owns the TMEM allocation (all 512 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(static_cast<uint32_t>(2 * kBwdBN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();                  // peers' barriers are initialised before anything is sent to them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ===== producer: 16 bulk copies per stage, one per lane =====
    // lane -> (operand, plane, chunk column of the k-block)
    const int op = (lane >> 3) & 1, pl = (lane >> 2) & 1, c = lane & 3;
    const __nv_bfloat16* plane = op == 0 ? (pl == 0 ? p.z_hi : p.z_lo) : (pl == 0 ? p.g_hi : p.g_lo);
    const int64_t plane_rows = op == 0 ? p.R_pad : p.D_pad;
    const uint32_t bytes = op == 0 ? kBM * 16u : kBwdBN * 16u;
    const uint32_t dst_off = (op == 0 ? pl * kTileA : 2 * kTileA + pl * kTileB) + static_cast<uint32_t>(c) * bytes;
    // the G tile's 8 copies are shared out over the cluster: this CTA issues the chunk columns c with c % CL == rank
    const bool mine = lane < 16 && (op == 0 || (c % CL) == rank);
    uint32_t it = 0;
    for (int w = cluster_id; w < p.total; w += num_clusters) {
      const int64_t m0 = (static_cast<int64_t>(w / p.tiles_n) * CL + rank) * kBM;
      const int64_t n0 = static_cast<int64_t>(w % p.tiles_n) * kBwdBN;
      const int64_t mn0 = op == 0 ? m0 : n0;
      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        const int s = static_cast<int>(it % kBwdStages);
        const uint32_t round = it / kBwdStages;
        if (round > 0) mbar_wait(&s_empty[s], (round - 1) & 1);
        if (lane == 0) mbar_arrive_expect_tx(&s_full[s], kBwdStageBytes);
        __syncwarp();
        if (mine) {
          const int64_t cc = static_cast<int64_t>(kb) * (kBK / 8) + c;
          uint8_t* dst = smem + static_cast<size_t>(s) * kBwdStageBytes + dst_off;
          const __nv_bfloat16* src = plane + (cc * plane_rows + mn0) * 8;
          if (CL > 1 && op == 1) tma_load_1d_multicast(dst, src, bytes, &s_full[s], kMask);
          else tma_load_1d(dst, src, bytes, &s_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(0, 0, kBwdBN);
      constexpr uint32_t csA = kBM * 16u, csB = kBwdBN * 16u;       // bytes between chunk columns (K-major tiles)
      uint32_t it = 0, local = 0;
      for (int w = cluster_id; w < p.total; w += num_clusters, ++local) {
        const uint32_t buf = local & 1, use = local >> 1;
        if (use > 0) {                                               // the epilogue must have drained this accumulator
          mbar_wait(&s_acc_empty[buf], (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t tmem_d = tmem_base + buf * kBwdBN;
        uint32_t accum = 0;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = static_cast<int>(it % kBwdStages);
          const uint32_t round = it / kBwdStages;
          mbar_wait(&s_full[s], round & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + static_cast<size_t>(s) * kBwdStageBytes);
          const uint32_t a_hi = st, a_lo = st + kTileA, b_hi = st + 2 * kTileA, b_lo = b_hi + kTileB;
#pragma unroll
          for (int term = 0; term < 3; ++term) {                     // hi*hi, hi*lo, lo*hi
            const uint32_t a_base = (term == 2) ? a_lo : a_hi;
            const uint32_t b_base = (term == 1) ? b_lo : b_hi;
#pragma unroll
            for (int ks = 0; ks < kBK / 16; ++ks) {
              umma_bf16(tmem_d, make_smem_desc(a_base + ks * 2u * csA, csA, 128u),
                        make_smem_desc(b_base + ks * 2u * csB, csB, 128u), idesc, accum);
              accum = 1;
            }
          }
          // stage reusable once these MMAs have read it -- in EVERY CTA whose producer writes into it
          if (CL > 1) umma_commit_multicast(&s_empty[s], kMask);
          else umma_commit(&s_empty[s]);
        }
        umma_commit(&s_acc_full[buf]);                               // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue =====
    const int q = warp & 3;                                          // TMEM lane quarter of this warp
    uint32_t local = 0;
    for (int w = cluster_id; w < p.total; w += num_clusters, ++local) {
      const int64_t m0 = (static_cast<int64_t>(w / p.tiles_n) * CL + rank) * kBM;
      const int n0 = (w % p.tiles_n) * kBwdBN;
      const uint32_t buf = local & 1, use = local >> 1;
      const int64_t r0 = m0 + q * 32, r = r0 + lane;                 // packed rows of this warp / lane
      const int t = static_cast<int>(r0 / p.Bp), b = static_cast<int>(r - static_cast<int64_t>(t) * p.Bp);
      const bool t_ok = t < p.Tn, row_ok = t_ok && b < p.Bn;
      float* orow = p.dz + (static_cast<int64_t>(b) * p.Tn + t) * p.D;
      float* prow = p.partial + (r0 >> 5) * p.D;
      // z~ of the statistics tensor for (r, n0+c .. +15): two chunk columns, hi + lo, 16-byte loads contiguous
      // across the 32 lanes.  They do not depend on the accumulator, and a chunk's four loads cost a full
      // DRAM/L2 round trip (the first version waited for them chunk by chunk: 35 % of all stall samples, the
      // epilogue -- not the tensor pipe -- set the tile time).  So kPre chunks are kept in flight in registers (this
      // kernel runs one CTA of 192 threads per SM: registers are plentiful), and the first kPre are issued BEFORE
      // the wait for the accumulator, i.e. under the main loop that is still producing it.
      constexpr int kPre = 8, kChunks = kBwdBN / 16;
      uint4 zq[kPre][4];
      auto z_load = [&](int slot, int c) {
        const int64_t cc = (n0 + c) >> 3;                              // always inside the padded planes
        zq[slot][0] = __ldg(reinterpret_cast<const uint4*>(p.s_hi + (cc * p.R_pad + r) * 8));
        zq[slot][1] = __ldg(reinterpret_cast<const uint4*>(p.s_lo + (cc * p.R_pad + r) * 8));
        zq[slot][2] = __ldg(reinterpret_cast<const uint4*>(p.s_hi + ((cc + 1) * p.R_pad + r) * 8));
        zq[slot][3] = __ldg(reinterpret_cast<const uint4*>(p.s_lo + ((cc + 1) * p.R_pad + r) * 8));
      };
#pragma unroll
      for (int i = 0; i < kPre; ++i) z_load(i, i * 16);
      mbar_wait(&s_acc_full[buf], use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * kBwdBN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
      for (int ci = 0; ci < kChunks; ++ci) {
        const int c = ci * 16;
        if (n0 + c >= p.D) break;
        const uint4 h0 = zq[ci % kPre][0], l0 = zq[ci % kPre][1], h1 = zq[ci % kPre][2], l1 = zq[ci % kPre][3];
        if (ci + kPre < kChunks) z_load(ci % kPre, c + kPre * 16);
        float v[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= p.alpha;
        if (row_ok) {                                                // D is a multiple of 8: whole float4s
#pragma unroll
          for (int h = 0; h < 4; ++h)
            if (n0 + c + 4 * h < p.D)
              *reinterpret_cast<float4*>(orow + n0 + c + 4 * h) = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
        }
        float z[16], zl[8];
        unpack_bf16x8(h0, z); unpack_bf16x8(l0, zl);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] += zl[i];
        unpack_bf16x8(h1, z + 8); unpack_bf16x8(l1, zl);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[8 + i] += zl[i];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= z[i];                   // padded rows: z~ = 0 and dz~ = 0
        // column sums over the warp's 32 rows: transposing butterfly, 16 -> 8 -> 4 -> 2 -> 1 values per lane
        float y8[8], y4[4], y2[2], y1;
        {
          const bool up = (lane & 16) != 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float keep = up ? v[8 + i] : v[i], send = up ? v[i] : v[8 + i];
            y8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
        }
        {
          const bool up = (lane & 8) != 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float keep = up ? y8[4 + i] : y8[i], send = up ? y8[i] : y8[4 + i];
            y4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
        }
        {
          const bool up = (lane & 4) != 0;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float keep = up ? y4[2 + i] : y4[i], send = up ? y4[i] : y4[2 + i];
            y2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
        }
        {
          const bool up = (lane & 2) != 0;
          const float keep = up ? y2[1] : y2[0], send = up ? y2[0] : y2[1];
          y1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        y1 += __shfl_xor_sync(0xffffffffu, y1, 1);
        const int col = n0 + c + (lane >> 1);                        // lane bits (4,3,2,1) = column bits (3,2,1,0)
        if (t_ok && (lane & 1) == 0 && col < p.D) prow[col] = y1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_acc_empty[buf]);
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();                  // no CTA leaves while a peer may still write into it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(2 * kBwdBN))
                 : "memory");
  }
}

}  // namespace

int umma_bwd_gemm(const PackedOperand& Z, const PackedOperand& G, const PackedOperand& Zs, const LfdRows& rows, int64_t D,
                  float alpha, float* dz, float* partial, cudaStream_t stream) {
  FDDM_CHECK_ARG(rows.tb_major, "umma_bwd_gemm: needs tb-major planes (B >= 32)");
  FDDM_CHECK_ARG(Z.hi && Z.lo && G.hi && G.lo && Zs.hi && Zs.lo && dz && partial, "umma_bwd_gemm: null pointer");
  FDDM_CHECK_ARG(D > 0 && D % 8 == 0, "umma_bwd_gemm: D must be a positive multiple of 8");
  FDDM_CHECK_ARG(Z.R_pad % kPackPad == 0 && Z.C_pad % kPackPad == 0 && G.R_pad == Z.C_pad && G.C_pad == Z.C_pad &&
                     Zs.R_pad == Z.R_pad && Zs.C_pad == Z.C_pad && Z.R_pad >= rows.rows_packed && Z.C_pad >= D,
                 "umma_bwd_gemm: packed operand padding");
  FDDM_CHECK_ARG(reinterpret_cast<uintptr_t>(dz) % 16 == 0, "umma_bwd_gemm: dz must be 16-byte aligned");
  BwdParams p;
  p.z_hi = Z.hi; p.z_lo = Z.lo; p.g_hi = G.hi; p.g_lo = G.lo; p.s_hi = Zs.hi; p.s_lo = Zs.lo;
  p.R_pad = Z.R_pad; p.D_pad = Z.C_pad;
  p.D = static_cast<int>(D);
  p.tiles_n = static_cast<int>((D + kBwdBN - 1) / kBwdBN);
  p.num_kb = static_cast<int>((D + kBK - 1) / kBK);
  constexpr int cl = 2;                                              // CTA pair sharing the multicast G tile
  int64_t tiles_m = (rows.rows_packed + kBM - 1) / kBM;              // R_pad is a multiple of 256 >= rows_packed
  tiles_m = (tiles_m + cl - 1) / cl * cl;                            // whole clusters (the extra tile is all zeros)
  const int64_t total = tiles_m / cl * p.tiles_n;                    // work items per cluster
  FDDM_CHECK_ARG(total < (1ll << 31) && tiles_m * kBM <= Z.R_pad, "umma_bwd_gemm: tile grid");
  p.total = static_cast<int>(total);
  p.Bn = rows.B; p.Bp = rows.Bp; p.Tn = rows.T;
  p.alpha = alpha; p.dz = dz; p.partial = partial;
  const size_t smem = static_cast<size_t>(kBwdStages) * kBwdStageBytes;
  const int clusters = static_cast<int>(std::min<int64_t>(total, num_sms() / cl));
  KernelScope ks("umma_bwd_persistent (z~ G + bn partials)", stream);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * cl));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cl);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FDDM_CUDA_OK(cudaFuncSetAttribute(umma_bwd_kernel<cl>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  FDDM_CUDA_OK(cudaLaunchKernelEx(&cfg, umma_bwd_kernel<cl>, p));
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

}  // namespace fddm
