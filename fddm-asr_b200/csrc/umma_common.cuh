// umma_common.cuh -- tcgen05 / TMEM / TMA helpers shared by the L_fd contraction kernels
// (lfd_umma.cu: forward split-K contraction; lfd_umma_bwd.cu: persistent backward contraction).
#pragma once

#include <cuda.h>      // CUtensorMap (types only; cuTensorMapEncodeTiled is fetched through the runtime)

#include "lfd_common.cuh"

namespace fddm {
namespace umma {

constexpr int kBM = 128;              // UMMA M (TMEM lanes)
constexpr int kBK = 32;               // K extent of one pipeline stage (two K=16 MMAs per term pair)
constexpr int kMaxBN = 256;
constexpr int kEpiWarps = 4;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr int kTmemCols = 256;
constexpr int kMaxStages = 8;
constexpr uint32_t kTileA = kBM * kBK * 2;       // bytes of one A plane per stage (8192)
constexpr uint32_t kTileB = kMaxBN * kBK * 2;    // bytes of one B plane per stage (16384)

// ---- tcgen05 / descriptor helpers ------------------------------------------------------------
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // SWIZZLE_NONE canonical layout; fields in 16-byte units.  bits [46,48) = descriptor version 1 (sm_100).
  return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_instr_desc(int a_mn_major, int b_mn_major, int n) {
  // kind::f16: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), majors at bits 15/16,
  // N>>3 at bits 17-22, M>>4 at bits 24-28.
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(kBM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 2-D tiled TMA load (SASS UTMALDG) completing on an mbarrier
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// number of bulk copies / bytes per copy that fill one plane of an operand tile
struct TilePlan {
  int ncopies;
  uint32_t bytes;          // per copy
  int64_t src_stride;      // elements between consecutive copies in the packed plane
  int64_t src_base;        // element offset of copy 0 for (mn0, k0)
};
__device__ __forceinline__ TilePlan plan_tile(const PackedOperand& op, int64_t mn0, int mn_extent, int64_t k0) {
  TilePlan t;
  if (op.mn_is_col) {      // chunk columns along MN, rows along K
    t.ncopies = mn_extent / 8;
    t.bytes = kBK * 16;
    t.src_stride = op.R_pad * 8;
    t.src_base = ((mn0 / 8) * op.R_pad + k0) * 8;
  } else {                 // chunk columns along K, rows along MN
    t.ncopies = kBK / 8;
    t.bytes = static_cast<uint32_t>(mn_extent) * 16;
    t.src_stride = op.R_pad * 8;
    t.src_base = ((k0 / 8) * op.R_pad + mn0) * 8;
  }
  return t;
}


}  // namespace umma
}  // namespace fddm
