// kl_kernels.cu -- fused diffusion-KL forward (+backward) over token rows.
//
// Replaces SchedulerAdapter.kl_term (reference train.py:190-255): softmax over the vocab, the two
// closed-form transition products, the un-normalised posteriors q and p_theta (+eps, quirk Q5), the
// per-token KL, the masked per-sample mean and the batch mean -- and, for training, d loss/d logits --
// in ONE pass over the logits (algorithmic traffic: s bytes/element forward, 2s forward+backward).
//
// Closed forms used per row (a_t = 1-beta_t, b_t = beta_t/K, a_p/b_p the same at t-1, beta_0 := 0;
// e_k = exp(z_k - max z), S = sum e, xh_k = e_k/S, eps = 1e-8):
//   P = b_t + a_t*xh[xt] + eps                      Q = b_t + a_t*[x0==xt] + eps
//   generic k (k != xt, k != x0):  q_g = b_t*b_p/Q,  p_k + eps = c1*e_k + c0,
//                                  c1 = b_t*a_p/(S*P), c0 = b_t*b_p/P + eps
//   KL_row   = q_g * sum_k (log(q_g+eps) - log(p_k+eps))  (+ exact corrections for k in {xt, x0})
//   backward : r_k = q_k/(p_k+eps),  S' = sum_k r_k p_k,  g_k = -r_k u_k a_p/P + [k=xt] a_t S'/P,
//              dKL/dz_n = xh_n (g_n - sum_m g_m xh_m)
// so the per-element work is one exp, one log (and one reciprocal for the gradient); the <=2 special
// entries are patched with their exact values.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "rowkit.cuh"

namespace fddm {
namespace {

struct KlWorkspace {
  unsigned int next_row;     // dynamic row scheduler (self-resetting)
  unsigned int done_ctas;    // completion counter (self-resetting)
  unsigned int pad[30];
  float kl_tok[1];           // [rows]
};

struct KlParams {
  const void* logits;
  const int64_t* xt;
  const int64_t* x0;
  const int64_t* t;
  const uint8_t* mask;       // [B, L] bytes (non-zero = valid), or
  const float* maskf;        // [B, L] fp32 weights -- the reference multiplies by x_mask.float() (train.py:250)
  const float* betas;
  const float* grad_scale;
  KlWorkspace* ws;
  float* loss_out;
  void* grad;
  int T, B, L, V, rows;
  float inv_bdiv;            // 1 / batch divisor
  int clamp_t;               // 1: clamp t into 1..T silently; 0: an out-of-range t traps (the reference raises)
  int batch;                 // rows a producer warp claims at once (<= 32, one lane per row's metadata)
};

constexpr float kEps = 1e-8f;
constexpr float kLn2 = 0.69314718055994530942f;

__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// weight of one token in the loss: mask/(sum_l mask + eps)/Bdiv, or 1/L/Bdiv without a mask
// (train.py:249-255).  Evaluated by one full warp.
// MASKF selects the fp32-weight variant at compile time: with it in the common (bool / no mask) instantiation the
// main kernel measured 9 % slower (199 vs 183 us at the c5 shard) although the code is never executed there.
template <bool MASKF>
__device__ __forceinline__ float token_weight_warp(const KlParams& p, int row, int lane) {
  const int b = row / p.L;
  if (MASKF) {
    const float* mrow = p.maskf + static_cast<size_t>(b) * p.L;
    float c = 0.0f;
    for (int l = lane; l < p.L; l += 32) c += mrow[l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return p.maskf[row] * (1.0f / (c + kEps)) * p.inv_bdiv;
  }
  if (p.mask == nullptr) return (1.0f / static_cast<float>(p.L)) * p.inv_bdiv;
  const uint8_t* mrow = p.mask + static_cast<size_t>(b) * p.L;
  int c = 0;
  for (int l = lane; l < p.L; l += 32) c += (mrow[l] != 0);
  c = warp_sum_int(c);
  const bool on = p.mask[row] != 0;
  return on ? (1.0f / (static_cast<float>(c) + kEps)) * p.inv_bdiv : 0.0f;
}

__device__ __forceinline__ void load_betas(const KlParams& p, int b, float& beta_t, float& beta_p) {
  long long tt = p.t[b];
  if ((tt < 1 || tt > p.T) && !p.clamp_t) {
    // the reference indexes betas[t-1]: IndexError on CPU, device-side assert on CUDA.  Same class of failure
    // here, without a host synchronisation or an extra launch on the hot path.
    printf("fddm kl_term: t[%d] = %lld is outside 1..%d\n", b, tt, p.T);
    __trap();
  }
  tt = tt < 1 ? 1 : (tt > p.T ? p.T : tt);
  beta_t = p.betas[tt - 1];
  beta_p = (tt == 1) ? 0.0f : p.betas[tt - 2];          // beta_0 := 0  (train.py:214-217)
}

// ------------------------------------------------------------------------------------------------
// the row math, shared by the register-resident and the shared-memory-resident paths
//
// Instruction budget: at 6.5 TB/s an fp32 forward+backward row stream leaves ~44 issue slots and ~5.5
// MUFU operations per element (a bf16 forward only a quarter of that), so logs and reciprocals are
// taken per GROUP OF 4 entries: with r_k = (p_k + eps) / (q_g + eps) (an O(1) quantity, so products of
// four stay far inside the fp32 range),
//   sum_k log2(q~/y_k)  = - sum_groups log2(r0 r1 r2 r3)                                  1 MUFU / 4
//   sum_k 1/r_k         =   sum_groups ((r0+r1) r2r3 + (r2+r3) r0r1) / (r0r1r2r3)         1 MUFU / 4
//   1/r_0               =   r1 * (r2r3) / (r0r1r2r3)   etc.  (gradient pass)              1 MUFU / 4
// ------------------------------------------------------------------------------------------------
template <int NT, bool BWD, typename T, class Row>
__device__ __forceinline__ float kl_row_math(Row& row, const int V, const int xt, const int x0, const float z_xt,
                                             const float z_x0, const float beta_t, const float beta_p,
                                             const float wscale, RedRing& red, T* grad_row) {
  constexpr float kLog2e = 1.4426950408889634f;
  const float Kf = static_cast<float>(V);
  const float a_t = 1.0f - beta_t, b_t = beta_t / Kf;
  const float a_p = 1.0f - beta_p, b_p = beta_p / Kf;

  // passes 1+2: per-thread max m_t and e_k = exp(z_k - m_t) kept in place, then ONE block reduction of the
  // (max, sum) pairs gives the row max m and S = sum_k exp(z_k - m).  The registers keep exp(z_k - m_t);
  // the per-thread factor f_t = exp(m_t - m) is folded into the coefficients of the later passes.
  float m_t = kNegInf;
  row.for_each_ro([&](int, float x) { m_t = fmaxf(m_t, x); });
  const float nm_t = -m_t * kLog2e;
  float s_t = 0.0f;
  row.for_each([&](int, float& x) {
    x = ex2_approx(fmaf(x, kLog2e, nm_t));
    s_t += x;
  });
  float m = m_t, S = s_t;
  block_softmax_stats<NT>(m, S, red);
  const float f_t = ex2_approx((m_t - m) * kLog2e);
  const float nm = -m * kLog2e;
  // Row scalars are evaluated redundantly by every thread, so they use MUFU reciprocals / logs
  // (~1e-7 relative, far inside the 1e-5 bar) instead of IEEE division / libm sequences.
  const float inv_S = rcp_approx(S);

  const bool same = (xt == x0);
  const float e_xt = ex2_approx(fmaf(z_xt, kLog2e, nm)), e_x0 = ex2_approx(fmaf(z_x0, kLog2e, nm));
  const float xh_xt = e_xt * inv_S, xh_x0 = e_x0 * inv_S;
  const float P = (b_t + a_t * xh_xt) + kEps;
  const float Q = (b_t + (same ? a_t : 0.0f)) + kEps;
  const float inv_P = rcp_approx(P), inv_Q = rcp_approx(Q);
  const float q_g = b_t * b_p * inv_Q;
  const float qt = q_g + kEps;                          // q~ = q_g + eps
  const float iq = rcp_approx(qt);
  const float c1 = b_t * a_p * inv_S * inv_P;           // p_k + eps = c1*e_k + c0   (generic k)
  const float c0e = b_t * b_p * inv_P;                  // c0 - eps
  const float c0 = c0e + kEps;
  const float c0r = c0 * iq;
  const float c1r = c1 * iq * f_t;                      // r_k = c1r*e_k + c0r with e_k relative to m_t
  const float cg = q_g * b_t * a_p * inv_P;

  // pass 3: generic terms for every k, four at a time
  //   a3[0] = sum log2(q~/y_k),  a3[1] = sum 1/r_k,  a3[2] = sum e_k/r_k
  float a3[BWD ? 3 : 1];
#pragma unroll
  for (int i = 0; i < (BWD ? 3 : 1); ++i) a3[i] = 0.0f;
  row.for_each4(
      [&](float* e) {
        const float r0 = fmaf(c1r, e[0], c0r), r1 = fmaf(c1r, e[1], c0r);
        const float r2 = fmaf(c1r, e[2], c0r), r3 = fmaf(c1r, e[3], c0r);
        const float p01 = r0 * r1, p23 = r2 * r3, pp = p01 * p23;
        a3[0] -= __log2f(pp);
        if (BWD) {
          const float ip = rcp_approx(pp);
          a3[1] = fmaf(fmaf(r0 + r1, p23, (r2 + r3) * p01), ip, a3[1]);
          a3[2] = fmaf(fmaf(fmaf(e[0], r1, e[1] * r0), p23, fmaf(e[2], r3, e[3] * r2) * p01), ip, a3[2]);
        }
      },
      [&](float& e) {
        const float r = fmaf(c1r, e, c0r);
        a3[0] -= __log2f(r);
        if (BWD) {
          const float ir = rcp_approx(r);
          a3[1] += ir;
          a3[2] = fmaf(e, ir, a3[2]);
        }
      });
  if (BWD) a3[2] *= f_t;                                // sum e_k/r_k in units of exp(z_k - m)
  block_sum<NT, (BWD ? 3 : 1)>(a3, red);
  // in y units: sum 1/y = iq * a3[1], sum e/y = iq * a3[2]
  const float sum_rc = BWD ? iq * a3[1] : 0.0f, sum_erc = BWD ? iq * a3[2] : 0.0f;

  // exact corrections for the special entries (computed redundantly by every thread)
  float kl = q_g * (kLn2 * a3[0]);
  // S' = sum_k r_k p_k over generic entries = q_g * sum (y_k - eps)/y_k = q_g * (c1*sum e/y + (c0-eps)*sum 1/y)
  float Sp = BWD ? q_g * fmaf(c1, sum_erc, c0e * sum_rc) : 0.0f;
  float G = BWD ? -cg * inv_S * sum_erc : 0.0f;
  float r_sp[2], u_sp[2], xh_sp[2];
  const int n_sp = same ? 1 : 2;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (i < n_sp) {
      const bool is_xt = (i == 0);
      const float e = is_xt ? e_xt : e_x0;
      const float xh = is_xt ? xh_xt : xh_x0;
      const float u = is_xt ? (b_t + a_t) : b_t;                       // train.py:227
      const float v = (same || !is_xt) ? (a_p + b_p) : b_p;            // train.py:230
      // what the generic formulas contributed for this k
      const float y = fmaf(c1, e, c0);
      const float rc = rcp_approx(y);
      kl += q_g * (kLn2 * __log2f(y * iq));
      // its true value
      const float qk = u * v * inv_Q;
      const float pk = u * (a_p * xh + b_p) * inv_P;
      const float rk = qk * rcp_approx(pk + kEps);
      kl += qk * (kLn2 * (__log2f(qk + kEps) - __log2f(pk + kEps)));
      if (BWD) {
        Sp += rk * pk - q_g * (y - kEps) * rc;
        G += cg * rc * xh;                                             // remove generic g_k*xh_k
        r_sp[i] = rk; u_sp[i] = u; xh_sp[i] = xh;
      }
    }
  }
  if (BWD) {
    float g_sp[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (i < n_sp) {
        g_sp[i] = -r_sp[i] * u_sp[i] * a_p * inv_P + (i == 0 ? a_t * Sp * inv_P : 0.0f);
        G += g_sp[i] * xh_sp[i];
      }
    }
    // pass 4: gradient  e_k * (A/y_k + Bc) = e_k * (Ar/r_k + Bc), generic entries from the row ...
    const float Ar = -wscale * inv_S * cg * iq * f_t;   // e_k here is exp(z_k - m_t): fold f_t in
    const float Bc = -wscale * inv_S * G * f_t;
    row.store4(
        grad_row,
        [&](const float* e, float* o) {
          const float r0 = fmaf(c1r, e[0], c0r), r1 = fmaf(c1r, e[1], c0r);
          const float r2 = fmaf(c1r, e[2], c0r), r3 = fmaf(c1r, e[3], c0r);
          const float p01 = r0 * r1, p23 = r2 * r3;
          const float ip = rcp_approx(p01 * p23);
          const float h01 = p23 * ip, h23 = p01 * ip;                  // 1/(r0 r1), 1/(r2 r3)
          o[0] = e[0] * fmaf(Ar, r1 * h01, Bc);
          o[1] = e[1] * fmaf(Ar, r0 * h01, Bc);
          o[2] = e[2] * fmaf(Ar, r3 * h23, Bc);
          o[3] = e[3] * fmaf(Ar, r2 * h23, Bc);
        },
        [&](float e) { return e * fmaf(Ar, rcp_approx(fmaf(c1r, e, c0r)), Bc); });
    // ... then the <= 2 special entries are overwritten with their exact values by the thread that owns them:
    // same thread, same address, program order -- no barrier needed
    if (row.owner_of(xt) == row.tid) Vec16<T>::store1(grad_row + xt, wscale * xh_sp[0] * (g_sp[0] - G));
    if (!same && row.owner_of(x0) == row.tid) Vec16<T>::store1(grad_row + x0, wscale * xh_sp[1] * (g_sp[1] - G));
  }
  return kl;
}

// last CTA: masked per-sample means and the batch mean, in a fixed order (deterministic)
template <int NT, bool MASKF>
__device__ __forceinline__ void kl_finalize(const KlParams& p, float* red, int tid) {
  constexpr int NW = NT / 32;
  const int warp = tid >> 5, lane = tid & 31;
  float acc = 0.0f;
  for (int b = warp; b < p.B; b += NW) {
    float s = 0.0f;
    int c = 0;
    float cf = 0.0f;
    for (int l = lane; l < p.L; l += 32) {
      const size_t i = static_cast<size_t>(b) * p.L + l;
      const float k = __ldcg(&p.ws->kl_tok[i]);
      if (MASKF) {
        const float wgt = p.maskf[i];
        cf += wgt;
        s += (wgt != 0.0f) ? k * wgt : 0.0f;
      } else if (p.mask) {
        const bool on = p.mask[i] != 0;
        c += on;
        s += on ? k : 0.0f;
      } else {
        s += k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    c = warp_sum_int(c);
    if (MASKF) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cf += __shfl_xor_sync(0xffffffffu, cf, o);
    }
    acc += MASKF ? s / (cf + kEps) : (p.mask ? s / (static_cast<float>(c) + kEps) : s / static_cast<float>(p.L));
  }
  if (lane == 0) red[warp] = acc;
  consumer_sync<NT>();
  if (tid == 0) {
    float tot = 0.0f;
    for (int w = 0; w < NW; ++w) tot += red[w];
    *p.loss_out = tot * p.inv_bdiv;
    p.ws->next_row = 0;
    p.ws->done_ctas = 0;
    __threadfence();
  }
}

template <int NT, bool MASKF>
__device__ __forceinline__ void kl_epilogue(const KlParams& p, float* red, int* s_flag, int tid) {
  consumer_sync<NT>();
  if (tid == 0) {
    __threadfence();
    const unsigned int d = atomicAdd(&p.ws->done_ctas, 1u);
    *s_flag = (d == gridDim.x - 1);
  }
  consumer_sync<NT>();
  if (*s_flag) {
    __threadfence();
    kl_finalize<NT, MASKF>(p, red, tid);
  }
}

// ------------------------------------------------------------------------------------------------
// fast path: TMA ring + register-resident rows
// ------------------------------------------------------------------------------------------------
template <typename T, int NT, int EPT, bool BWD, bool MASKF>
__global__ void __launch_bounds__(NT + 32, (NT <= 256 ? 2 : 1))
kl_rows_ring_kernel(const KlParams p, const int nstages, const uint32_t stage_bytes) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ uint64_t s_full[kMaxStages], s_empty[kMaxStages];
  __shared__ RingMeta s_meta[kMaxStages];
  __shared__ float s_red[kRedFloats];
  __shared__ int s_flag;

  Ring ring;
  ring.stages = dyn_smem;
  ring.stage_bytes = stage_bytes;
  ring.nstages = nstages;
  ring.full = s_full;
  ring.empty = s_empty;
  ring.meta = s_meta;
  ring_init<NT>(ring);

  const int tid = threadIdx.x;
  const uint32_t row_bytes = static_cast<uint32_t>(p.V) * sizeof(T);

  if (tid >= NT) {
    // ===== producer warp =====
    // A row's metadata costs a chain of dependent global loads (row claim -> mask row -> ids, t -> betas), about
    // 2 us, and a single lane doing that row after row bounded the whole kernel (the forward-only kernel took
    // as long per row as forward+backward: the consumers were starved).  So the warp claims `batch` consecutive
    // rows at once and every lane prepares ONE row's metadata in parallel; the copies are then issued in row
    // order as stages free up, with nothing but the stage hand-shake on that path.
    const int lane = tid - NT;
    const int batch = p.batch;
    int s = 0;
    uint32_t round = 0;
    for (;;) {
      int base = 0;
      if (lane == 0) base = static_cast<int>(atomicAdd(&p.ws->next_row, static_cast<unsigned int>(batch)));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= p.rows) {
        if (round > 0) mbar_wait_backoff(&ring.empty[s], (round - 1) & 1);
        if (lane == 0) {
          ring.meta[s].row = -1;
          mbar_arrive(&ring.full[s]);
        }
        break;
      }
      const int nb = min(batch, p.rows - base);
      const int row = base + lane;                       // this lane's row (lanes >= nb idle)
      const bool mine = lane < nb;
      const int b = (mine ? row : base) / p.L;
      // token weights: one warp-cooperative count per distinct sample of the batch
      float w = 0.0f;
      if (!MASKF && p.mask == nullptr) {
        w = (1.0f / static_cast<float>(p.L)) * p.inv_bdiv;
      } else {
        const int b_first = base / p.L, b_last = (base + nb - 1) / p.L;
        for (int bb = b_first; bb <= b_last; ++bb) {
          float inv;
          if (MASKF) {
            const float* mrow = p.maskf + static_cast<size_t>(bb) * p.L;
            float c = 0.0f;
            for (int l = lane; l < p.L; l += 32) c += mrow[l];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            inv = 1.0f / (c + kEps);
          } else {
            const uint8_t* mrow = p.mask + static_cast<size_t>(bb) * p.L;
            int c = 0;
            for (int l = lane; l < p.L; l += 32) c += (mrow[l] != 0);
            inv = 1.0f / (static_cast<float>(warp_sum_int(c)) + kEps);
          }
          if (mine && b == bb) w = (MASKF ? p.maskf[row] : (p.mask[row] != 0 ? 1.0f : 0.0f)) * inv * p.inv_bdiv;
        }
      }
      RingMeta mt;
      mt.row = row;
      mt.w = w;
      mt.i0 = mt.i1 = 0;
      mt.f0 = mt.f1 = mt.f2 = mt.f3 = 0.0f;
      mt.r0 = mt.r1 = mt.r2 = 0u;
      if (mine) {
        mt.i0 = static_cast<int>(p.xt[row]);
        mt.i1 = static_cast<int>(p.x0[row]);
        load_betas(p, b, mt.f0, mt.f1);
      }
      for (int k = 0; k < nb; ++k) {
        if (round > 0) mbar_wait_backoff(&ring.empty[s], (round - 1) & 1);
        if (lane == k) {
          ring.meta[s] = mt;
          if (w != 0.0f) {
            mbar_arrive_expect_tx(&ring.full[s], row_bytes);
            tma_load_1d(ring.stage(s), static_cast<const uint8_t*>(p.logits) + static_cast<size_t>(row) * row_bytes,
                        row_bytes, &ring.full[s]);
          } else {
            mbar_arrive(&ring.full[s]);
          }
        }
        __syncwarp();
        if (++s == nstages) { s = 0; ++round; }
      }
    }
    return;
  }

  // ===== consumers =====
  const float gscale = (BWD && p.grad_scale) ? __ldg(p.grad_scale) : 1.0f;
  RedRing red{s_red, 0};
  RegRow<T, NT, EPT> row;
  row.tid = tid;
  row.nvec = p.V / RegRow<T, NT, EPT>::N;
  int s = 0;
  uint32_t round = 0;
  for (;;) {
    mbar_wait(&ring.full[s], round & 1);
    const RingMeta mt = ring.meta[s];
    if (mt.row < 0) break;
    T* grad_row = BWD ? static_cast<T*>(p.grad) + static_cast<size_t>(mt.row) * p.V : nullptr;
    if (mt.w == 0.0f) {
      ring_release(ring, s);
      if (BWD) row.store(grad_row, [](int, float) { return 0.0f; });
      if (tid == 0) p.ws->kl_tok[mt.row] = 0.0f;
    } else {
      const T* st = reinterpret_cast<const T*>(ring.stage(s));
      row.load_from_smem(st, p.V, tid);
      const float z_xt = Vec16<T>::load1(st + mt.i0);
      const float z_x0 = Vec16<T>::load1(st + mt.i1);
      ring_release(ring, s);
      const float kl = kl_row_math<NT, BWD, T>(row, p.V, mt.i0, mt.i1, z_xt, z_x0, mt.f0, mt.f1, mt.w * gscale, red,
                                               grad_row);
      if (tid == 0) p.ws->kl_tok[mt.row] = kl;
    }
    if (++s == nstages) { s = 0; ++round; }
  }
  kl_epilogue<NT, MASKF>(p, s_red, &s_flag, tid);
}

// ------------------------------------------------------------------------------------------------
// generic path: any V <= FDDM_MAX_VOCAB, any alignment; the row is an fp32 copy in shared memory
// ------------------------------------------------------------------------------------------------
template <typename T, int NT, bool BWD, bool MASKF>
__global__ void __launch_bounds__(NT, 1) kl_rows_generic_kernel(const KlParams p) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ float s_red[kRedFloats];
  __shared__ int s_flag;
  float* srow = reinterpret_cast<float*>(dyn_smem);
  const int tid = threadIdx.x;
  const float gscale = (BWD && p.grad_scale) ? __ldg(p.grad_scale) : 1.0f;
  RedRing red{s_red, 0};
  SmemRow<T, NT> row;
  for (int r = blockIdx.x; r < p.rows; r += gridDim.x) {
    const float w = token_weight_warp<MASKF>(p, r, tid & 31);
    T* grad_row = BWD ? static_cast<T*>(p.grad) + static_cast<size_t>(r) * p.V : nullptr;
    if (w == 0.0f) {
      if (BWD) for (int k = tid; k < p.V; k += NT) Vec16<T>::store1(grad_row + k, 0.0f);
      if (tid == 0) p.ws->kl_tok[r] = 0.0f;
      continue;
    }
    const T* src = static_cast<const T*>(p.logits) + static_cast<size_t>(r) * p.V;
    row.load_from_gmem(srow, src, p.V, tid);
    const int xt = static_cast<int>(p.xt[r]), x0 = static_cast<int>(p.x0[r]);
    const float z_xt = srow[xt], z_x0 = srow[x0];
    float beta_t, beta_p;
    load_betas(p, r / p.L, beta_t, beta_p);
    consumer_sync<NT>();      // everyone has read the specials before pass 2 overwrites the row
    const float kl = kl_row_math<NT, BWD, T>(row, p.V, xt, x0, z_xt, z_x0, beta_t, beta_p, w * gscale, red, grad_row);
    if (tid == 0) p.ws->kl_tok[r] = kl;
    consumer_sync<NT>();      // row buffer is reused by the next iteration
  }
  kl_epilogue<NT, MASKF>(p, s_red, &s_flag, tid);
}

// x *= num/den, all CTAs leave immediately when the ratio is exactly 1
template <typename T>
__global__ void __launch_bounds__(256) scale_inplace_kernel(T* x, int64_t nvec, int64_t n, const float* num,
                                                            const float* den) {
  const float r = __ldg(num) / (den ? __ldg(den) : 1.0f);
  if (r == 1.0f) return;
  constexpr int N = Vec16<T>::N;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    uint4 v = *(reinterpret_cast<const uint4*>(x) + i);
    float f[N];
    Vec16<T>::unpack(v, f);
#pragma unroll
    for (int e = 0; e < N; ++e) f[e] *= r;
    *(reinterpret_cast<uint4*>(x) + i) = Vec16<T>::pack(f);
  }
  // tail (n not a multiple of the vector width)
  if (blockIdx.x == 0) {
    for (int64_t k = nvec * N + threadIdx.x; k < n; k += blockDim.x) Vec16<T>::store1(x + k, Vec16<T>::load1(x + k) * r);
  }
}

template <typename T, bool BWD>
int launch_kl(const KlParams& p, cudaStream_t stream) {
  const size_t row_bytes = static_cast<size_t>(p.V) * sizeof(T);
  const bool aligned = (row_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(p.logits) % 16 == 0) &&
                       (!BWD || reinterpret_cast<uintptr_t>(p.grad) % 16 == 0);
  const int sms = row_kernel_sms();
  KernelScope ks(BWD ? "kl_rows_fwdbwd" : "kl_rows_fwd", stream);
  if (aligned && p.V <= 32768) {
    int nt, ept;
    if (p.V <= 4096) { nt = 128; ept = 32; }
    else if (p.V <= 8192) { nt = 256; ept = 32; }
    else if (p.V <= 16384) { nt = 512; ept = 32; }
    else { nt = 512; ept = 64; }
    const RingPlan plan = plan_ring(row_bytes, nt);
    if (plan.nstages >= 1) {
      const int grid = static_cast<int>(std::min<int64_t>(p.rows, static_cast<int64_t>(sms) * plan.ctas_per_sm));
      const uint32_t sb = static_cast<uint32_t>((row_bytes + 127) & ~size_t(127));
      // rows per claim: the largest power of two that still leaves every CTA >= 16 claims (load balance: measured
      // at 16384 rows, 2 rows per claim is the best and 32 costs 25%), at most one row per producer lane
      KlParams pb = p;
      pb.batch = 1;
      while (pb.batch < 32 && static_cast<int64_t>(pb.batch) * 2 * grid * 16 <= p.rows) pb.batch *= 2;
#define FDDM_KL_LAUNCH(NT_, EPT_)                                                                           \
  do {                                                                                                      \
    auto kfn = p.maskf ? kl_rows_ring_kernel<T, NT_, EPT_, BWD, true> : kl_rows_ring_kernel<T, NT_, EPT_, BWD, false>; \
    FDDM_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                                      static_cast<int>(plan.smem_bytes)));                                  \
    kfn<<<grid, NT_ + 32, plan.smem_bytes, stream>>>(pb, plan.nstages, sb);                                 \
  } while (0)
      if (nt == 128) FDDM_KL_LAUNCH(128, 32);
      else if (nt == 256) FDDM_KL_LAUNCH(256, 32);
      else if (ept == 32) FDDM_KL_LAUNCH(512, 32);
      else FDDM_KL_LAUNCH(512, 64);
#undef FDDM_KL_LAUNCH
      FDDM_LAUNCH_OK();
      return FDDM_OK;
    }
  }
  // generic path
  const size_t smem = static_cast<size_t>(p.V) * sizeof(float) + 128;
  auto kfn = p.maskf ? kl_rows_generic_kernel<T, 256, BWD, true> : kl_rows_generic_kernel<T, 256, BWD, false>;
  FDDM_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int grid = static_cast<int>(std::min<int64_t>(p.rows, static_cast<int64_t>(sms) * 4));
  kfn<<<grid, 256, smem, stream>>>(p);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int kl_entry(const void* logits, int dtype, const int64_t* xt, const int64_t* x0, const int64_t* t,
             const void* x_mask, int flags, const float* betas, int64_t T, int64_t B, int64_t L, int64_t V, double batch_div,
             const float* grad_scale, void* workspace, float* loss_out, void* grad_logits, bool bwd,
             cudaStream_t stream) {
  FDDM_CHECK_ARG(logits && xt && x0 && t && betas && workspace && loss_out, "kl: null pointer argument");
  FDDM_CHECK_ARG(!bwd || grad_logits, "kl: grad_logits is null");
  FDDM_CHECK_ARG(dtype_valid(dtype), "kl: bad dtype %d", dtype);
  FDDM_CHECK_ARG(B > 0 && L > 0 && V > 1 && T > 0, "kl: non-positive size B=%lld L=%lld V=%lld T=%lld", (long long)B,
                 (long long)L, (long long)V, (long long)T);
  FDDM_CHECK_ARG(B * L < (1ll << 31), "kl: too many rows");
  FDDM_CHECK_ARG(batch_div > 0.0, "kl: batch_div must be positive");
  if (V > FDDM_MAX_VOCAB) {
    set_error("kl: V=%lld exceeds FDDM_MAX_VOCAB=%d", (long long)V, FDDM_MAX_VOCAB);
    return FDDM_EUNSUPPORTED;
  }
  KlParams p;
  p.logits = logits; p.xt = xt; p.x0 = x0; p.t = t; p.betas = betas;
  const bool mask_is_f32 = (flags & FDDM_KL_MASK_F32) != 0;
  p.mask = mask_is_f32 ? nullptr : static_cast<const uint8_t*>(x_mask);
  p.maskf = mask_is_f32 ? static_cast<const float*>(x_mask) : nullptr;
  p.clamp_t = (flags & FDDM_KL_CLAMP_T) ? 1 : 0;
  p.batch = 1;
  p.grad_scale = grad_scale; p.ws = static_cast<KlWorkspace*>(workspace); p.loss_out = loss_out;
  p.grad = grad_logits;
  p.T = static_cast<int>(T); p.B = static_cast<int>(B); p.L = static_cast<int>(L); p.V = static_cast<int>(V);
  p.rows = static_cast<int>(B * L);
  p.inv_bdiv = static_cast<float>(1.0 / batch_div);
  if (dtype == FDDM_F32) return bwd ? launch_kl<float, true>(p, stream) : launch_kl<float, false>(p, stream);
  if (dtype == FDDM_BF16)
    return bwd ? launch_kl<__nv_bfloat16, true>(p, stream) : launch_kl<__nv_bfloat16, false>(p, stream);
  return bwd ? launch_kl<__half, true>(p, stream) : launch_kl<__half, false>(p, stream);
}

}  // namespace
}  // namespace fddm

extern "C" {

size_t fddm_kl_workspace_bytes(int64_t B, int64_t L) {
  if (B <= 0 || L <= 0) return 0;
  return 128 + static_cast<size_t>(B) * static_cast<size_t>(L) * sizeof(float);
}

int fddm_kl_forward(const void* logits, int dtype, const int64_t* xt, const int64_t* x0, const int64_t* t,
                    const void* x_mask, int flags, const float* betas, int64_t T, int64_t B, int64_t L, int64_t V,
                    double batch_div, void* workspace, float* loss_out, fddm_stream_t stream) {
  FDDM_API_RANGE();
  return fddm::kl_entry(logits, dtype, xt, x0, t, x_mask, flags, betas, T, B, L, V, batch_div, nullptr, workspace,
                        loss_out, nullptr, false, reinterpret_cast<cudaStream_t>(stream));
}

int fddm_kl_forward_backward(const void* logits, int dtype, const int64_t* xt, const int64_t* x0, const int64_t* t,
                             const void* x_mask, int flags, const float* betas, int64_t T, int64_t B, int64_t L,
                             int64_t V,
                             double batch_div, const float* grad_scale, void* workspace, float* loss_out,
                             void* grad_logits, fddm_stream_t stream) {
  FDDM_API_RANGE();
  return fddm::kl_entry(logits, dtype, xt, x0, t, x_mask, flags, betas, T, B, L, V, batch_div, grad_scale,
                        workspace, loss_out, grad_logits, true, reinterpret_cast<cudaStream_t>(stream));
}

int fddm_scale_inplace(void* x, int dtype, int64_t n, const float* num, const float* den, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(x && num, "scale_inplace: null pointer");
  FDDM_CHECK_ARG(dtype_valid(dtype), "scale_inplace: bad dtype %d", dtype);
  FDDM_CHECK_ARG(n >= 0, "scale_inplace: negative size");
  FDDM_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0, "scale_inplace: x must be 16-byte aligned");
  if (n == 0) return FDDM_OK;
  const int grid = num_sms() * 8;
  KernelScope ks("scale_inplace_kernel", stream);
  if (dtype == FDDM_F32) {
    scale_inplace_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(x), n / 4, n, num, den);
  } else if (dtype == FDDM_BF16) {
    scale_inplace_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(x), n / 8, n, num, den);
  } else {
    scale_inplace_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<__half*>(x), n / 8, n, num, den);
  }
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

}  // extern "C"
