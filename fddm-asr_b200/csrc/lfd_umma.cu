// lfd_umma.cu -- the tcgen05 (5th-gen tensor core) contraction used by L_fd.
//
//   out[s][m][n] = alpha * sum_{k in split s} A(m,k) * B(n,k)          fp32 accumulation in TMEM
//
// Operands are row-major global matrices that are *transformed on the way into shared memory*: each
// producer thread loads 8 contiguous elements (one 16-byte bf16 chunk after conversion), applies the
// per-(position, channel) standardisation x~ = x*rstd - mean*rstd of losses/fddm_losses.py:23-26,
// splits the fp32 value into bf16 hi + bf16 residual (so that hi*hi + hi*lo + lo*hi carries ~16
// mantissa bits through the bf16 tensor pipe; 1e-5 parity needs more than TF32's 10) and stores both
// into UMMA's canonical no-swizzle layout.  So z_a / z_b are read once per tile straight from
// HBM/L2 -- no standardised copy is ever materialised -- and no TMA tensor map is involved.
//
// Shared-memory operand tile (per term): R rows x C columns of the global matrix (C contiguous),
// stored as 16-byte chunks:   chunk(cc, r)  at  cc*CS + (r/8)*128 + (r%8)*16,  CS = (R/8)*128 + 32.
// A "core matrix" = 8 rows x 16 bytes = 128 contiguous bytes.  The +32 pad makes the producers'
// stores conflict-free (a quarter-warp writes 4 chunk-columns x 2 rows = 8 distinct 16-byte slots)
// while their global loads stay row-contiguous (4 lanes x 32 B = one 128-byte line per row).
//   * operand read "down the rows" (MN index = column, K index = row; the forward's z~^T z~):
//     MN-major descriptor, SBO (stride between 8-element MN chunks) = CS, LBO (between 8-row K groups) = 128
//   * operand read "along the rows" (MN index = row, K index = column; the backward's z~ G):
//     K-major descriptor, SBO (between 8-row MN groups) = 128, LBO (between 8-element K chunks) = CS
//
// CTA = 8 producer/epilogue warps + 1 MMA warp.  mbarrier ring: producers -> full[s] (256 arrivals
// after fence.proxy.async) -> one elected thread issues tcgen05.mma kind::f16 (M=128, N<=256, K=16)
// -> tcgen05.commit -> empty[s].  The last commit signals the epilogue, which reads the 128 x N
// fp32 accumulator out of TMEM with tcgen05.ld (32 lanes x 16 columns per instruction).
#include <algorithm>

#include "lfd_common.cuh"

namespace fddm {
namespace {

constexpr int kBM = 128;              // UMMA M (TMEM lanes)
constexpr int kBK = 32;               // K extent of one pipeline stage (two K=16 MMAs per term pair)
constexpr int kMaxBN = 256;
constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kThreads = kProducerThreads + 32;
constexpr int kTmemCols = 256;
constexpr int kMaxStages = 6;

__host__ __device__ constexpr uint32_t chunk_stride(int R) { return static_cast<uint32_t>(R / 8) * 128u + 32u; }
// bytes of one operand tile (one term) for either orientation, upper bound over both
constexpr uint32_t kTileA = 16 * chunk_stride(kBK);       // MN-major: 128/8 chunk columns of 32 rows = 8704
constexpr uint32_t kTileB = 32 * chunk_stride(kBK);       // 256/8 chunk columns                      = 17408
static_assert(4 * chunk_stride(kBM) <= kTileA, "K-major A tile must fit");
static_assert(4 * chunk_stride(kMaxBN) <= kTileB, "K-major B tile must fit");

struct GemmParams {
  UmmaOperand A, B;
  int M, N, K;
  int BN;                 // UMMA N of this launch (multiple of 32, <= 256)
  int tiles_m, tiles_n, splits;
  int k_per_split;        // multiple of kBK
  int terms, stages;
  float alpha;
  float* out;
  int64_t out_ld, out_split_stride;
};

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

// ---- producer: one operand tile, global -> registers -> (standardise, split) -> shared ---------
template <int TERMS>
__device__ __forceinline__ void load_operand_tile(const UmmaOperand& op, int64_t mn0, int mn_extent, int64_t k0,
                                                  int64_t k_end, uint8_t* s_hi, uint8_t* s_lo, int warp, int lane) {
  const int R = op.mn_is_col ? kBK : mn_extent;
  const int C = op.mn_is_col ? mn_extent : kBK;
  const int64_t row0 = op.mn_is_col ? k0 : mn0;
  const int64_t col0 = op.mn_is_col ? mn0 : k0;
  const int64_t row_end = op.mn_is_col ? min(op.nrows, k_end) : op.nrows;
  const int64_t col_end = op.mn_is_col ? op.ncols : min(op.ncols, k_end);
  const uint32_t CS = chunk_stride(R);
  const int ncb = C / 32;                        // 32-column blocks (4 chunks) across the tile
  const int nwb = ncb * (R / 8);                 // warp-blocks: 8 rows x 4 chunks
  for (int wb = warp; wb < nwb; wb += kProducerWarps) {
    const int cb = wb % ncb, rb = wb / ncb;
    const int cc = cb * 4 + (lane & 3);
    const int r = rb * 8 + (lane >> 2);
    const int64_t gr = row0 + r, gc = col0 + static_cast<int64_t>(cc) * 8;
    float x[8];
    if (gr < row_end && gc < col_end) {          // extents are multiples of 8 columns: all-or-nothing
      if (op.dtype == FDDM_F32) {
        const float* p = static_cast<const float*>(op.ptr) + gr * op.ld + gc;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
      } else if (op.dtype == FDDM_BF16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(op.ptr) + gr * op.ld + gc));
        Vec16<__nv_bfloat16>::unpack(v, x);
      } else {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(static_cast<const __half*>(op.ptr) + gr * op.ld + gc));
        Vec16<__half>::unpack(v, x);
      }
      if (op.scale != nullptr) {
        const int64_t so = (gr % op.T) * op.stat_ld + gc;
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(op.scale + so));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(op.scale + so) + 1);
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(op.shift + so));
        const float4 h1 = __ldg(reinterpret_cast<const float4*>(op.shift + so) + 1);
        x[0] = fmaf(x[0], s0.x, h0.x); x[1] = fmaf(x[1], s0.y, h0.y);
        x[2] = fmaf(x[2], s0.z, h0.z); x[3] = fmaf(x[3], s0.w, h0.w);
        x[4] = fmaf(x[4], s1.x, h1.x); x[5] = fmaf(x[5], s1.y, h1.y);
        x[6] = fmaf(x[6], s1.z, h1.z); x[7] = fmaf(x[7], s1.w, h1.w);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = 0.0f;
    }
    const uint32_t off = static_cast<uint32_t>(cc) * CS + static_cast<uint32_t>(r >> 3) * 128u +
                         static_cast<uint32_t>(r & 7) * 16u;
    const uint4 hi = Vec16<__nv_bfloat16>::pack(x);
    *reinterpret_cast<uint4*>(s_hi + off) = hi;
    if (TERMS == 2) {
      float h[8];
      Vec16<__nv_bfloat16>::unpack(hi, h);
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = x[e] - h[e];          // exact in fp32
      *reinterpret_cast<uint4*>(s_lo + off) = Vec16<__nv_bfloat16>::pack(h);
    }
  }
}

template <int TERMS>
__global__ void __launch_bounds__(kThreads, 1) umma_gemm_kernel(const GemmParams p) {
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
  const int num_kb = static_cast<int>((k_end - k_begin + kBK - 1) / kBK);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&s_full[s], kProducerThreads);
      mbar_init(&s_empty[s], 1);
    }
    mbar_init(&s_accum, 1);
    mbar_fence_init();
  }
  if (warp == kProducerWarps) {                      // the MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = s_tmem_base;

  if (warp < kProducerWarps) {
    // ===== producers =====
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % p.stages;
      const uint32_t round = static_cast<uint32_t>(kb / p.stages);
      if (round > 0) mbar_wait(&s_empty[s], (round - 1) & 1);
      uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
      uint8_t* a_hi = st;
      uint8_t* a_lo = st + kTileA;                   // only with TERMS == 2
      uint8_t* b_hi = st + TERMS * kTileA;
      uint8_t* b_lo = b_hi + kTileB;
      const int64_t k0 = k_begin + static_cast<int64_t>(kb) * kBK;
      load_operand_tile<TERMS>(p.A, m0, kBM, k0, k_end, a_hi, a_lo, warp, lane);
      load_operand_tile<TERMS>(p.B, n0, p.BN, k0, k_end, b_hi, b_lo, warp, lane);
      fence_proxy_async();                           // generic-proxy stores -> visible to the MMA (async proxy)
      mbar_arrive(&s_full[s]);
    }
    // ===== epilogue: TMEM -> registers -> global =====
    mbar_wait(&s_accum, 0);
    tc_fence_after();
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int64_t m = m0 + q * 32 + lane;
    float* orow = p.out + static_cast<int64_t>(split) * p.out_split_stride + m * p.out_ld;
    for (int c = (warp >> 2) * 16; c < p.BN; c += 32) {
      float v[16];
      tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), v);
      const int64_t n = n0 + c;
      if (m < p.M && n < p.N) {                      // N is a multiple of 8; handle the 16-column chunk in halves
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
  } else {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      const uint32_t idesc = make_instr_desc(p.A.mn_is_col, p.B.mn_is_col, p.BN);
      const uint32_t csA = chunk_stride(p.A.mn_is_col ? kBK : kBM);
      const uint32_t csB = chunk_stride(p.B.mn_is_col ? kBK : p.BN);
      const uint32_t lboA = p.A.mn_is_col ? 128u : csA, sboA = p.A.mn_is_col ? csA : 128u;
      const uint32_t lboB = p.B.mn_is_col ? 128u : csB, sboB = p.B.mn_is_col ? csB : 128u;
      const uint32_t advA = p.A.mn_is_col ? 256u : 2u * csA;     // K += 16 within a stage
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
          const uint32_t a_base = (term == 2) ? a_lo : a_hi;     // hi*hi, hi*lo, lo*hi
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
  }
  __syncthreads();
  if (warp == kProducerWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                 "r"(static_cast<uint32_t>(kTmemCols))
                 : "memory");
  }
}

bool operand_ok(const UmmaOperand& o) {
  return o.ptr && reinterpret_cast<uintptr_t>(o.ptr) % 16 == 0 && o.ld % 8 == 0 && o.ncols % 8 == 0 &&
         (o.scale == nullptr || (o.shift != nullptr && reinterpret_cast<uintptr_t>(o.scale) % 16 == 0 &&
                                 reinterpret_cast<uintptr_t>(o.shift) % 16 == 0 && o.stat_ld % 4 == 0 && o.T > 0));
}

}  // namespace

int umma_gemm(const UmmaOperand& A, const UmmaOperand& B, int64_t M, int64_t N, int64_t K, int splits, int terms,
              float alpha, float* out, int64_t out_ld, int64_t out_split_stride, cudaStream_t stream) {
  FDDM_CHECK_ARG(operand_ok(A) && operand_ok(B), "umma_gemm: operand alignment (16-byte pointers, ld %% 8 == 0)");
  FDDM_CHECK_ARG(M > 0 && N > 0 && K > 0 && splits >= 1 && (terms == 1 || terms == 2), "umma_gemm: bad shape");
  FDDM_CHECK_ARG(out && reinterpret_cast<uintptr_t>(out) % 16 == 0 && out_ld % 4 == 0 && N % 4 == 0 &&
                     out_split_stride % 4 == 0,
                 "umma_gemm: output alignment");
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
  if (terms == 2) {
    FDDM_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    umma_gemm_kernel<2><<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(p);
  } else {
    FDDM_CUDA_OK(cudaFuncSetAttribute(umma_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    umma_gemm_kernel<1><<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(p);
  }
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

}  // namespace fddm
