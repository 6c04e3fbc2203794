// lfd_umma.cu -- the tcgen05 (5th-gen tensor core) contraction used by L_fd.
//
//   out[s][m][n] = alpha * sum_{k in split s} A(m,k) * B(n,k)          fp32 accumulation in TMEM
//
// Operands arrive in the packed bf16 (hi, lo) planes described in lfd_common.cuh: every run of rows of
// one 8-column chunk is already a column of UMMA core matrices, so a pipeline stage is filled by a
// handful of 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) completing on an mbarrier:
//
//   warp 0      producer: lane 0 arms full[s] with the stage's byte count, all 32 lanes issue copies
//   warp 1      MMA issuer: one elected lane, tcgen05.mma cta_group::1 kind::f16, M=128, N<=256, K=16;
//               tcgen05.commit -> empty[s] frees the stage, the last commit signals the epilogue
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 16 columns) TMEM -> registers -> global
//
// fp32 accuracy: hi*hi + hi*lo + lo*hi (three MMAs per K step) carries ~16 mantissa bits through the
// bf16 pipe; 1e-5 parity on correlated inputs needs more than TF32's 10.
//
// Shared-memory tile (one plane, dense): chunk column cc, row r at  cc*(R*16) + r*16  (R = rows in the
// tile), i.e. 8-row x 16-byte core matrices of 128 contiguous bytes.
//   MN index = column (forward):  R = BK rows;  MN-major descriptor, SBO = BK*16, LBO = 128
//   MN index = row    (backward): R = TM rows;  K-major descriptor,  SBO = 128,   LBO = TM*16
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "umma_common.cuh"

namespace fddm {
namespace {

struct GemmParams {
  PackedOperand A, B;
  int M, N, K;
  int BN;                 // UMMA N of this launch (multiple of 32, <= 256)
  int tiles_m, tiles_n, splits;
  int k_per_split;        // multiple of kBK
  int terms, stages;
  float alpha;
  float* out;
  int64_t out_ld, out_split_stride;
};

using namespace umma;

// The four tensor maps describe the (hi, lo) planes of A and B as 2-D arrays [C_pad/8][R_pad*8]: an
// MN-major tile (BK rows x TM/8 chunk columns) is then ONE box {BK*8 elements, TM/8} that lands densely
// in the canonical layout.  They are only dereferenced for operands with mn_is_col == 1; K-major tiles
// are a few long contiguous runs and use 1-D bulk copies.
template <int TERMS>
__global__ void __launch_bounds__(kThreads, 1)
umma_gemm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t s_full[kMaxStages], s_empty[kMaxStages], s_accum;
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_bytes = TERMS * (kTileA + kTileB);

  // tile / split of this CTA (CTAs of one split are adjacent: they share operand rows in L2)
  const int tiles = p.tiles_m * p.tiles_n;
  const int split = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int tm = tile / p.tiles_n, tn = tile % p.tiles_n;
  const int64_t m0 = static_cast<int64_t>(tm) * kBM, n0 = static_cast<int64_t>(tn) * p.BN;
  const int64_t k_begin = static_cast<int64_t>(split) * p.k_per_split;
  const int64_t k_end = min(static_cast<int64_t>(p.K), k_begin + p.k_per_split);
  const int num_kb = static_cast<int>((k_end - k_begin + kBK - 1) / kBK);   // padded planes hold zeros past K

  if (tid == 0) {
    if (p.A.mn_is_col) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_hi) : "memory");
      if (TERMS == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA_lo) : "memory");
    }
    if (p.B.mn_is_col) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_hi) : "memory");
      if (TERMS == 2) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB_lo) : "memory");
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    mbar_init(&s_accum, 1);
    mbar_fence_init();
  }
  if (warp == 1) {                                   // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = s_tmem_base;

  if (warp == 0) {
    // ===== producer: 1-D TMA bulk copies of packed core-matrix columns =====
    const uint32_t bytesA = kBM * kBK * 2, bytesB = static_cast<uint32_t>(p.BN) * kBK * 2;
    const uint32_t tx = TERMS * (bytesA + bytesB);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % p.stages;
      const uint32_t round = static_cast<uint32_t>(kb / p.stages);
      if (round > 0) mbar_wait(&s_empty[s], (round - 1) & 1);
      if (lane == 0) mbar_arrive_expect_tx(&s_full[s], tx);
      __syncwarp();
      uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
      const int64_t k0 = k_begin + static_cast<int64_t>(kb) * kBK;
      const TilePlan ta = plan_tile(p.A, m0, kBM, k0), tb = plan_tile(p.B, n0, p.BN, k0);
#pragma unroll
      for (int term = 0; term < TERMS; ++term) {
        const __nv_bfloat16* pa = term == 0 ? p.A.hi : p.A.lo;
        const __nv_bfloat16* pb = term == 0 ? p.B.hi : p.B.lo;
        uint8_t* da = st + term * kTileA;
        uint8_t* db = st + TERMS * kTileA + term * kTileB;
        if (p.A.mn_is_col) {
          if (lane == 0) tma_load_2d(da, term == 0 ? &tmA_hi : &tmA_lo, static_cast<int>(k0 * 8), static_cast<int>(m0 / 8), &s_full[s]);
        } else {
          for (int c = lane; c < ta.ncopies; c += 32)
            tma_load_1d(da + static_cast<size_t>(c) * ta.bytes, pa + ta.src_base + c * ta.src_stride, ta.bytes, &s_full[s]);
        }
        if (p.B.mn_is_col) {
          if (lane == 1) tma_load_2d(db, term == 0 ? &tmB_hi : &tmB_lo, static_cast<int>(k0 * 8), static_cast<int>(n0 / 8), &s_full[s]);
        } else {
          for (int c = lane; c < tb.ncopies; c += 32)
            tma_load_1d(db + static_cast<size_t>(c) * tb.bytes, pb + tb.src_base + c * tb.src_stride, tb.bytes, &s_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(p.A.mn_is_col, p.B.mn_is_col, p.BN);
      const uint32_t csA = p.A.mn_is_col ? kBK * 16u : kBM * 16u;          // bytes between chunk columns
      const uint32_t csB = p.B.mn_is_col ? kBK * 16u : static_cast<uint32_t>(p.BN) * 16u;
      const uint32_t lboA = p.A.mn_is_col ? 128u : csA, sboA = p.A.mn_is_col ? csA : 128u;
      const uint32_t lboB = p.B.mn_is_col ? 128u : csB, sboB = p.B.mn_is_col ? csB : 128u;
      const uint32_t advA = p.A.mn_is_col ? 256u : 2u * csA;               // K += 16 within a stage
      const uint32_t advB = p.B.mn_is_col ? 256u : 2u * csB;
      uint32_t accum = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % p.stages;
        const uint32_t round = static_cast<uint32_t>(kb / p.stages);
        mbar_wait(&s_full[s], round & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
        const uint32_t a_hi = st, a_lo = st + kTileA, b_hi = st + TERMS * kTileA, b_lo = b_hi + kTileB;
#pragma unroll
        for (int term = 0; term < (TERMS == 2 ? 3 : 1); ++term) {
          const uint32_t a_base = (term == 2) ? a_lo : a_hi;               // hi*hi, hi*lo, lo*hi
          const uint32_t b_base = (term == 1) ? b_lo : b_hi;
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks) {
            const uint64_t da = make_smem_desc(a_base + ks * advA, lboA, sboA);
            const uint64_t db = make_smem_desc(b_base + ks * advB, lboB, sboB);
            umma_bf16(tmem_d, da, db, idesc, accum);
            accum = 1;
          }
        }
        umma_commit(&s_empty[s]);                    // stage reusable once these MMAs have read it
      }
      umma_commit(&s_accum);                         // accumulator complete
    }
    __syncwarp();
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    mbar_wait(&s_accum, 0);
    tc_fence_after();
    const int q = warp & 3;                          // TMEM lane quarter this warp may access (warp id % 4)
    const int64_t m = m0 + q * 32 + lane;
    float* orow = p.out + static_cast<int64_t>(split) * p.out_split_stride + m * p.out_ld;
    for (int c = 0; c < p.BN; c += 16) {
      float v[16];
      tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
      const int64_t n = n0 + c;
      if (m < p.M && n < p.N) {                      // N is a multiple of 4: 16-byte stores
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if (n + 4 * h < p.N) {
            float4 o = make_float4(v[4 * h] * p.alpha, v[4 * h + 1] * p.alpha, v[4 * h + 2] * p.alpha,
                                   v[4 * h + 3] * p.alpha);
            *reinterpret_cast<float4*>(orow + n + 4 * h) = o;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
  }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// plane [C_pad/8][R_pad*8] bf16, box {BK*8, mn_extent/8}, no swizzle, no interleave
int make_plane_map(CUtensorMap* tm, const __nv_bfloat16* plane, int64_t R_pad, int64_t C_pad, int mn_extent) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) {
    set_error("umma_gemm: cuTensorMapEncodeTiled is not available from the driver");
    return FDDM_ECUDA;
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(R_pad) * 8, static_cast<cuuint64_t>(C_pad / 8)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(R_pad) * 16};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK * 8), static_cast<cuuint32_t>(mn_extent / 8)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(plane), gdim, gstride, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("umma_gemm: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return FDDM_ECUDA;
  }
  return FDDM_OK;
}

bool operand_ok(const PackedOperand& o, int terms) {
  return o.hi && (terms == 1 || o.lo) && reinterpret_cast<uintptr_t>(o.hi) % 16 == 0 &&
         reinterpret_cast<uintptr_t>(o.lo) % 16 == 0 && o.R_pad % kPackPad == 0 && o.C_pad % kPackPad == 0;
}

}  // namespace

int umma_gemm(const PackedOperand& A, const PackedOperand& B, int64_t M, int64_t N, int64_t K, int splits, int terms,
              float alpha, float* out, int64_t out_ld, int64_t out_split_stride, cudaStream_t stream) {
  FDDM_CHECK_ARG(terms == 1 || terms == 2, "umma_gemm: terms must be 1 or 2");
  FDDM_CHECK_ARG(operand_ok(A, terms) && operand_ok(B, terms), "umma_gemm: packed operand alignment / padding");
  FDDM_CHECK_ARG(M > 0 && N > 0 && K > 0 && splits >= 1, "umma_gemm: bad shape");
  FDDM_CHECK_ARG(out && reinterpret_cast<uintptr_t>(out) % 16 == 0 && out_ld % 4 == 0 && N % 4 == 0 &&
                     out_split_stride % 4 == 0,
                 "umma_gemm: output alignment");
  // padded extents must cover every tile the grid touches
  const int64_t mA = A.mn_is_col ? A.C_pad : A.R_pad, kA = A.mn_is_col ? A.R_pad : A.C_pad;
  const int64_t nB = B.mn_is_col ? B.C_pad : B.R_pad, kB = B.mn_is_col ? B.R_pad : B.C_pad;
  GemmParams p;
  p.A = A; p.B = B;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.BN = static_cast<int>(std::min<int64_t>(kMaxBN, (N + 31) / 32 * 32));
  p.tiles_m = static_cast<int>((M + kBM - 1) / kBM);
  p.tiles_n = static_cast<int>((N + p.BN - 1) / p.BN);
  int64_t kps = (K + splits - 1) / splits;
  kps = (kps + kBK - 1) / kBK * kBK;
  p.k_per_split = static_cast<int>(kps);
  p.splits = static_cast<int>((K + kps - 1) / kps);         // every split owns >= 1 k-block
  FDDM_CHECK_ARG(p.splits <= splits, "umma_gemm: internal split error");
  const int64_t k_cover = (K + kBK - 1) / kBK * kBK;
  FDDM_CHECK_ARG(static_cast<int64_t>(p.tiles_m) * kBM <= mA && static_cast<int64_t>(p.tiles_n) * p.BN <= nB &&
                     k_cover <= kA && k_cover <= kB,
                 "umma_gemm: packed planes do not cover the tile grid");
  p.terms = terms;
  const uint32_t stage_bytes = static_cast<uint32_t>(terms) * (kTileA + kTileB);
  p.stages = std::min<int>(kMaxStages, static_cast<int>((200u * 1024u) / stage_bytes));
  p.alpha = alpha; p.out = out; p.out_ld = out_ld; p.out_split_stride = out_split_stride;
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes;
  const int64_t grid = static_cast<int64_t>(p.tiles_m) * p.tiles_n * p.splits;
  FDDM_CHECK_ARG(grid < (1ll << 31), "umma_gemm: grid too large");
  // splits that the rounding removed must still be defined for the fixed-order reduction
  if (p.splits < splits && out_split_stride > 0)
    FDDM_CUDA_OK(cudaMemsetAsync(out + static_cast<int64_t>(p.splits) * out_split_stride, 0,
                                 sizeof(float) * static_cast<size_t>(splits - p.splits) * out_split_stride, stream));
  CUtensorMap tm[4];
  memset(tm, 0, sizeof(tm));
  if (A.mn_is_col) {
    if (int rc = make_plane_map(&tm[0], A.hi, A.R_pad, A.C_pad, kBM)) return rc;
    if (terms == 2) if (int rc = make_plane_map(&tm[1], A.lo, A.R_pad, A.C_pad, kBM)) return rc;
  }
  if (B.mn_is_col) {
    if (int rc = make_plane_map(&tm[2], B.hi, B.R_pad, B.C_pad, p.BN)) return rc;
    if (terms == 2) if (int rc = make_plane_map(&tm[3], B.lo, B.R_pad, B.C_pad, p.BN)) return rc;
  }
  KernelScope ks(A.mn_is_col ? "umma_gemm_fwd (z~^T z~, split-K)" : "umma_gemm_bwd (z~ G)", stream);
  if (terms == 2) {
    FDDM_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    umma_gemm_kernel<2><<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(p, tm[0], tm[1], tm[2], tm[3]);
  } else {
    FDDM_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    umma_gemm_kernel<1><<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(p, tm[0], tm[1], tm[2], tm[3]);
  }
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

}  // namespace fddm
