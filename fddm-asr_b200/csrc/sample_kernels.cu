// sample_kernels.cu -- the two resampling kernels of the token path.
//
//  * sample_q_ids  : SchedulerAdapter.sample_q (reference train.py:180-188) fused ids -> ids.  The
//    one-hot [B,L,V] tensor, q_sample's five passes over it and torch.multinomial's exponential_/
//    div/argmax are replaced by one streaming pass over the injected Exp(1) noise (4 B/element) or
//    by in-kernel Philox (no memory traffic at all).  For a one-hot x0 the row of q_sample takes two
//    values (p_hi at x0, p_lo elsewhere, sched:44-49), so the kernel evaluates
//    argmax_k fdiv(p_k, E_k) with the reference's fp32 roundings and lowest-index tie rule.
//  * jump_step     : DiffusionJumpySampler._jump_once minus the decoder (sampler:189-215): softmax,
//    Delta-step posterior for a one-hot x_t (sched:186-206) or alpha-bar mix (sampler:139-151), then
//    argmax / exponential-race sampling (== Categorical.sample(), sampler:153-162).
//
// Elementwise steps that decide ids use __fmul_rn/__fadd_rn/__fdiv_rn so that nvcc cannot contract
// them into FMAs: the op order is the reference's.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "rowkit.cuh"

namespace fddm {
namespace {

// ------------------------------------------------------------------------------------------------
// noise sources
// ------------------------------------------------------------------------------------------------
struct NoiseMem {           // injected Exp(1) variates for one row (shared or global memory)
  const float* p;
  bool vec_ok;              // row pointer 16-byte aligned
  __device__ __forceinline__ void load4(int k0, float* E) const {
    if (vec_ok) {
      const float4 v = *reinterpret_cast<const float4*>(p + k0);
      E[0] = v.x; E[1] = v.y; E[2] = v.z; E[3] = v.w;
    } else {
      E[0] = p[k0]; E[1] = p[k0 + 1]; E[2] = p[k0 + 2]; E[3] = p[k0 + 3];
    }
  }
  __device__ __forceinline__ float load1(int k) const { return p[k]; }
};

struct NoisePhilox {        // counter = (k/4 within the row, row, offset); lane k%4 of the output
  uint2 key;
  uint32_t row;
  uint2 off;
  __device__ __forceinline__ void load4(int k0, float* E) const {
    const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(k0 >> 2), row, off.x, off.y), key);
    E[0] = exp1_from_bits(r.x); E[1] = exp1_from_bits(r.y);
    E[2] = exp1_from_bits(r.z); E[3] = exp1_from_bits(r.w);
  }
  __device__ __forceinline__ float load1(int k) const {
    float E[4];
    load4(k & ~3, E);
    return E[k & 3];
  }
};

// (seed, offset) either from the launch parameters or, when a device pointer is given, from device
// memory -- so that a captured CUDA graph draws fresh noise on every replay
__device__ __forceinline__ void philox_key_off(const uint64_t* state, uint2 key_in, uint2 off_in, uint2& key, uint2& off) {
  key = key_in; off = off_in;
  if (state != nullptr) {
    const uint64_t sd = state[0], of = state[1];
    key = make_uint2(static_cast<uint32_t>(sd), static_cast<uint32_t>(sd >> 32));
    off = make_uint2(static_cast<uint32_t>(of), static_cast<uint32_t>(of >> 32));
  }
}

// f(k, x&, E_k) over a RegRow / SmemRow
template <typename T, int NT, int EPT, class Noise, class F>
__device__ __forceinline__ void for_each_noise(RegRow<T, NT, EPT>& row, const Noise& nz, F&& f) {
  constexpr int N = RegRow<T, NT, EPT>::N;
#pragma unroll
  for (int j = 0; j < RegRow<T, NT, EPT>::NVEC; ++j) {
    const int vi = j * NT + row.tid;
    if (vi < row.nvec) {
      float E[N];
#pragma unroll
      for (int q = 0; q < N / 4; ++q) nz.load4(vi * N + 4 * q, E + 4 * q);
#pragma unroll
      for (int e = 0; e < N; ++e) f(vi * N + e, row.v[j * N + e], E[e]);
    }
  }
}
template <typename T, int NT, class Noise, class F>
__device__ __forceinline__ void for_each_noise(SmemRow<T, NT>& row, const Noise& nz, F&& f) {
  row.for_each([&](int k, float& x) { f(k, x, nz.load1(k)); });
}

#ifndef FDDM_JUMP_DT
// ------------------------------------------------------------------------------------------------
// sample_q_ids
// ------------------------------------------------------------------------------------------------
struct SampleQParams {
  const int64_t* x0;
  const int64_t* t;
  const float* alpha_bar;
  const float* noise;
  int64_t* out;
  int T, L, K, rows;
  float eps, u;              // u = fp32(1/K)  (sched:45)
  uint2 key, off;
  const uint64_t* philox_state;   // optional device {seed, offset}: overrides key/off (CUDA-graph replay)
};

// the two values of the q_sample row for one-hot x0, with the reference's roundings (sched:44-49)
__device__ __forceinline__ void q_sample_two_values(float ab, float u, float eps, int K, float& p_hi, float& p_lo) {
  const float lo_raw = __fmul_rn(__fsub_rn(1.0f, ab), u);          // ab*0 + (1-ab)*u
  const float hi_raw = __fadd_rn(ab, lo_raw);                       // ab*1 + (1-ab)*u
  const float hi_c = fmaxf(hi_raw, eps), lo_c = fmaxf(lo_raw, eps); // clamp_min(eps), quirk Q4
  // row sum: one p_hi and K-1 p_lo; evaluated in fp64 then rounded (torch's fp32 tree differs by <=1ulp)
  const float s = fmaxf(static_cast<float>(static_cast<double>(hi_c) + static_cast<double>(K - 1) * lo_c), eps);
  p_hi = __fdiv_rn(hi_c, s);
  p_lo = __fdiv_rn(lo_c, s);
}

template <int NT, bool PHILOX>
__global__ void __launch_bounds__(NT) sample_q_kernel(const SampleQParams p) {
  __shared__ float s_red[kRedFloats];
  RedRing red{s_red, 0};
  const int tid = threadIdx.x;
  const bool vec = (p.K % 4 == 0) && (PHILOX || reinterpret_cast<uintptr_t>(p.noise) % 16 == 0);
  for (int row = blockIdx.x; row < p.rows; row += gridDim.x) {
    const int b = row / p.L;
    long long tt = p.t[b];
    tt = tt < 1 ? 1 : (tt > p.T ? p.T : tt);
    const float ab = p.alpha_bar[tt - 1];
    float p_hi, p_lo;
    q_sample_two_values(ab, p.u, p.eps, p.K, p_hi, p_lo);
    const int x0 = static_cast<int>(p.x0[row]);
    float best = -1.0f;
    int best_k = 0x7fffffff;
    auto visit = [&](int k, float E) {
      const float sc = __fdiv_rn(k == x0 ? p_hi : p_lo, E);
      if (sc > best) { best = sc; best_k = k; }                    // strict: earlier index keeps ties
    };
    NoisePhilox ph;
    philox_key_off(p.philox_state, p.key, p.off, ph.key, ph.off); ph.row = static_cast<uint32_t>(row);
    const float* nrow = PHILOX ? nullptr : p.noise + static_cast<size_t>(row) * p.K;
    if (vec) {
      const int nvec = p.K / 4;
      constexpr int U = 4;
      int vi = tid;
      for (; vi + (U - 1) * NT < nvec; vi += U * NT) {
        float E[U][4];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          if (PHILOX) {
            ph.load4((vi + j * NT) * 4, E[j]);
          } else {
            const uint4 v = ldg_stream_v4(nrow + (vi + j * NT) * 4);
            E[j][0] = __uint_as_float(v.x); E[j][1] = __uint_as_float(v.y);
            E[j][2] = __uint_as_float(v.z); E[j][3] = __uint_as_float(v.w);
          }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) visit((vi + j * NT) * 4 + e, E[j][e]);
        }
      }
      for (; vi < nvec; vi += NT) {
        float E[4];
        if (PHILOX) {
          ph.load4(vi * 4, E);
        } else {
          const uint4 v = ldg_stream_v4(nrow + vi * 4);
          E[0] = __uint_as_float(v.x); E[1] = __uint_as_float(v.y);
          E[2] = __uint_as_float(v.z); E[3] = __uint_as_float(v.w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) visit(vi * 4 + e, E[e]);
      }
    } else {
      for (int k = tid; k < p.K; k += NT) visit(k, PHILOX ? ph.load1(k) : nrow[k]);
    }
    block_argmax<NT>(best, best_k, red);
    if (tid == 0) p.out[row] = best_k;
  }
}


// In-kernel RNG path: for a one-hot x0 the categorical q(x_t|x_0) has two values, so a draw needs O(1)
// work per token instead of a race over all K entries: keep x0 with probability p_hi, otherwise pick
// one of the other K-1 ids uniformly.  Same distribution as torch.multinomial on the q_sample row
// (the per-entry race is only needed to replay *injected* noise bit for bit, see sample_q_kernel).
__global__ void __launch_bounds__(256) sample_q_closed_kernel(const SampleQParams p) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.rows) return;
  long long tt = p.t[row / p.L];
  tt = tt < 1 ? 1 : (tt > p.T ? p.T : tt);
  float p_hi, p_lo;
  q_sample_two_values(p.alpha_bar[tt - 1], p.u, p.eps, p.K, p_hi, p_lo);
  const int x0 = static_cast<int>(p.x0[row]);
  uint2 key, off;
  philox_key_off(p.philox_state, p.key, p.off, key, off);
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(row), 0u, off.x, off.y), key);
  const float u1 = static_cast<float>(r.x >> 8) * (1.0f / 16777216.0f);             // [0,1)
  int out = x0;
  if (!(u1 < p_hi) && p.K > 1) {
    // uniform over the K-1 other ids: 32 random bits scaled by (K-1) without modulo bias worth noting
    int j = static_cast<int>(__umulhi(r.y, static_cast<uint32_t>(p.K - 1)));
    out = j + (j >= x0 ? 1 : 0);
  }
  p.out[row] = out;
}

#endif  // !FDDM_JUMP_DT

// ------------------------------------------------------------------------------------------------
// jump_step
// ------------------------------------------------------------------------------------------------
}  // namespace

// Shared by the per-dtype translation units (this file is compiled once per logits dtype with
// -DFDDM_JUMP_DT=<n> for the kernels, and once without it for the C entry points).
struct JumpParams {
  const void* logits;
  const int64_t* x_t;
  const float* coeffs;       // exact: a_cum[B] | b_cum[B] | a_tgt[B] | b_tgt[B] | identity
  const float* alpha_bar;    // fast
  const float* noise;
  int64_t* x_out;
  int64_t* argmax_p_out;     // optional: argmax_k p_x0 (the sampler's final x_0, sampler:292)
  void* p_out;
  unsigned int* work;        // self-resetting counters: [0] next row, [1] CTAs done, [2] rows on the fallback list
  int* row_list;             // fallback list (greedy streamed kernel -> bit-faithful kernel), [rows] ints after the counters
  int list_mode;             // 1: rows are row_list[0 .. work[2]) instead of 0 .. rows
  // streamed kernels: a row is copied in `nchunks` pieces of `chunk_vecs` 16-byte vectors (a multiple of the
  // consumer thread count, so a thread owns the same vector residues in every chunk) through `nstages` stages
  int chunk_vecs, nchunks, nstages;
  int B, L, K, rows;
  int flags;
  int abar_index;
  float temperature, eps, u;
  uint2 key, off;
  const uint64_t* philox_state;   // optional device {seed, offset}: overrides key/off (CUDA-graph replay)
};
int jump_launch_f32(const JumpParams& p, int noise, cudaStream_t stream);
int jump_launch_bf16(const JumpParams& p, int noise, cudaStream_t stream);
int jump_launch_f16(const JumpParams& p, int noise, cudaStream_t stream);

#ifdef FDDM_JUMP_DT
namespace {

// counter-domain tag of the jump kernels (xor-ed into counter word 3) so that a jump and a sample_q
// call sharing one {seed, offset} never evaluate Philox on the same counter
constexpr uint32_t kJumpDomain = 0x4A554D50u;

struct JumpRowCtx {
  uint32_t r0, r1, r2;       // this row's random bits (drawn once per row by the producer)
  int xt;
  float a_c, b_c, a_g, b_g;  // exact
  float ab;                  // fast
  bool identity;
};

// The production sampling path: in-kernel RNG, temperature 1.  ONE block barrier per row.
//   * per thread: m_t = max z, e_k = exp(z_k - m_t) kept in registers, s_t = sum e_k; per warp (shuffles):
//     (m_w, s_w) with s_w relative to m_w; lane 0 publishes the pair; barrier; every thread derives the row's
//     (m, S) from the NW pairs.  p_k = e_k * sc_t with sc_t = exp(m_t - m) / S.
//   * the un-normalised target weights are affine in p_k (exact: w_k = A_k (a_g p_k + b_g sum p),
//     fast: w_k = abar p_k + (1-abar)/K), so the probability mass of a whole WARP follows from its s_w, its
//     element count and -- for the warp that owns x_t -- one correction: every thread knows all NW warp masses
//     right after that one barrier, without any further communication.
//   * hierarchical inverse-CDF draw with three uniforms from ONE Philox call keyed by the row (all threads
//     evaluate the same call, so they agree without talking): pick the warp from the NW warp masses; the other
//     warps are done with the row; inside the picked warp pick the lane from the lane masses (shuffle prefix
//     sum), then the entry among that lane's registers (gathered across the warp by shuffles, prefix sum,
//     ballot).  The picked lane writes the id.  Same distribution as Categorical(probs=w/sum w).sample().
// `wr` != nullptr (the last jump of a chain: p_x0 and its argmax are wanted) adds a pass that rounds p to the
// logits dtype (quirk Q11), stores it, and one more block reduction for the argmax.
template <int NT, typename T, class Row>
__device__ __forceinline__ void jump_row_fast(Row& row, const JumpParams& p, const JumpRowCtx& c, const float z_xt, RedRing& red, const int* s_nw, T* p_row_out,
                                              int64_t* x_out_row, int64_t* argmax_out_row) {
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int NW = NT / 32;
  const bool exact = (p.flags & FDDM_JUMP_EXACT) != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // four independent accumulators: the chains are EPT/4 long instead of EPT
  float mx[4] = {kNegInf, kNegInf, kNegInf, kNegInf};
  row.for_each4([&](float* x) { mx[0] = fmaxf(mx[0], x[0]); mx[1] = fmaxf(mx[1], x[1]);
                                mx[2] = fmaxf(mx[2], x[2]); mx[3] = fmaxf(mx[3], x[3]); },
                [&](float&) {});
  const float m_t = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
  const float nm_t = -m_t * kLog2e;
  float sx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  row.for_each4([&](float* x) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { x[e] = ex2_approx(fmaf(x[e], kLog2e, nm_t)); sx[e] += x[e]; }
  }, [&](float&) {});
  const float s_t = (sx[0] + sx[1]) + (sx[2] + sx[3]);
  // warp-level (max, sum): max first, then the rescaled sums
  float m_w = m_t;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, o));
  const float f_tw = ex2_approx((m_t - m_w) * kLog2e);          // e_k are relative to m_t: rescale to m_w
  float s_w = s_t * f_tw;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s_w += __shfl_xor_sync(0xffffffffu, s_w, o);
  float* sc = red.next();
  if (lane == 0) { sc[warp] = m_w; sc[32 + warp] = s_w; }
  consumer_sync<NT>();
  float m = sc[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) m = fmaxf(m, sc[w]);
  float sw[NW];                                                 // warp sums relative to the row max
  float S = 0.0f;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    sw[w] = sc[32 + w] * ex2_approx((sc[w] - m) * kLog2e);
    S += sw[w];
  }
  const float inv_S = rcp_approx(S);
  float rs = f_tw * ex2_approx((m_w - m) * kLog2e) * inv_S;      // p_k = (register) * rs
  bool rounded = false;                                          // registers hold p rounded to T (rs == 1)

  if (p_row_out != nullptr || argmax_out_row != nullptr) {
    float pm = -1.0f;
    int pm_k = 0x7fffffff;
    row.for_each([&](int k, float& x) {
      x = Vec16<T>::round_trip(x * rs);
      if (x > pm) { pm = x; pm_k = k; }
    });
    rs = 1.0f;
    rounded = true;
    if (p_row_out != nullptr) row.store(p_row_out, [](int, float x) { return x; });
    if (argmax_out_row != nullptr) {
      block_argmax<NT>(pm, pm_k, red);
      if (threadIdx.x == 0) *argmax_out_row = pm_k;
    }
  }
  if (c.identity) {                                              // sched:133-134 (delta <= 0)
    if (threadIdx.x == 0) *x_out_row = c.xt;
    if (p.flags & FDDM_JUMP_DEBUG_W) {
      T* w_row = static_cast<T*>(p.p_out) + static_cast<size_t>(x_out_row - p.x_out) * p.K;
      row.store(w_row, [&](int k, float) { return k == c.xt ? 1.0f : 0.0f; });
    }
    return;
  }

  // w_k = wa * p_k + wb, except (wa_x, wb_x) at k == x_t
  const float p_xt = Vec16<T>::round_trip(ex2_approx(fmaf(z_xt, kLog2e, -m * kLog2e)) * inv_S);
  float wa, wb, wa_x, wb_x;
  if (exact) {
    const float bs = c.b_g;                                      // b_tgt * sum(x0hat), sum = 1     sched:191
    wa = c.b_c * c.a_g; wb = c.b_c * bs;                         // A = b_cum * sum_xt             sched:187
    wa_x = (c.a_c + c.b_c) * c.a_g; wb_x = (c.a_c + c.b_c) * bs; // A = a_cum + b_cum at x_t
  } else {
    wa = wa_x = c.ab;
    wb = wb_x = (1.0f - c.ab) * p.u;                             // sampler:147-151
  }
  const float corr = fmaf(wa_x - wa, p_xt, wb_x - wb);           // extra mass of the entry k == x_t
  const int owner = row.owner_of(c.xt);                          // thread that holds x_t
  constexpr float k2m24 = 1.0f / 16777216.0f;
  // level 1: the warp
  float tot = 0.0f;
  float cw[NW];
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    float mw = fmaf(wa, sw[w] * inv_S, wb * static_cast<float>(s_nw[w]));
    if (w == (owner >> 5)) mw += corr;
    tot += fmaxf(mw, 0.0f);
    cw[w] = tot;
  }
  if (p.flags & FDDM_JUMP_DEBUG_W) {                            // test hook: the normalised target distribution
    T* w_row = static_cast<T*>(p.p_out) + static_cast<size_t>(x_out_row - p.x_out) * p.K;
    const float inv_tot = 1.0f / tot;
    row.store(w_row, [&](int k, float x) {
      const float pk = x * rs;
      return fmaxf((k == c.xt) ? fmaf(wa_x, pk, wb_x) : fmaf(wa, pk, wb), 0.0f) * inv_tot;
    });
  }
  const float t1 = tot * ((static_cast<float>(c.r0 >> 8) + 0.5f) * k2m24);
  int wsel = NW - 1;
#pragma unroll
  for (int w = NW - 1; w >= 0; --w)
    if (cw[w] >= t1 && (w == 0 ? cw[0] > 0.0f : cw[w] > cw[w - 1])) wsel = w;
  if (warp != wsel) return;

  // level 2: the lane (inclusive prefix sum of the lane masses)
  float psum = s_t * rs;
  if (rounded) {
    psum = 0.0f;
    row.for_each([&](int, float& x) { psum += x; });
  }
  float mass = fmaf(wa, psum, wb * static_cast<float>(row.n_owned()));
  if (row.tid == owner) mass += corr;
  mass = fmaxf(mass, 0.0f);
  float pre = mass;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, pre, o);
    if (lane >= o) pre += up;
  }
  const float wtot = __shfl_sync(0xffffffffu, pre, 31);
  const float t2 = wtot * ((static_cast<float>(c.r1 >> 8) + 0.5f) * k2m24);
  unsigned hit = __ballot_sync(0xffffffffu, pre >= t2 && mass > 0.0f);
  if (hit == 0u) hit = __ballot_sync(0xffffffffu, mass > 0.0f);   // rounding pushed the target past the total
  const int lsel = hit ? (__ffs(hit) - 1) : 0;

  // level 3: the entry among the picked lane's registers, 32 at a time across the warp
  const float rs_sel = __shfl_sync(0xffffffffu, rs, lsel);
  const int tid_sel = (warp << 5) + lsel;
  const float lane_mass = __shfl_sync(0xffffffffu, mass, lsel);
  const float t3 = lane_mass * ((static_cast<float>(c.r2 >> 8) + 0.5f) * k2m24);
  constexpr int N = Row::N, EPT = Row::NVEC * Row::N;
  float base = 0.0f;
  int chosen = -1, last_pos = -1;
#pragma unroll
  for (int r0 = 0; r0 < EPT; r0 += 32) {
    float mine = 0.0f;
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float v = __shfl_sync(0xffffffffu, row.v[r0 + r], lsel);
      if (lane == r) mine = v;
    }
    const int reg = r0 + lane;
    const int vi = (reg / N) * NT + tid_sel;
    const int k = vi * N + (reg % N);
    const bool valid = vi < row.nvec;
    float w = 0.0f;
    if (valid) w = (k == c.xt) ? fmaf(wa_x, mine * rs_sel, wb_x) : fmaf(wa, mine * rs_sel, wb);
    w = fmaxf(w, 0.0f);
    float pw = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, pw, o);
      if (lane >= o) pw += up;
    }
    pw += base;
    const unsigned pos = __ballot_sync(0xffffffffu, w > 0.0f);
    const unsigned got = __ballot_sync(0xffffffffu, w > 0.0f && pw >= t3);
    if (chosen < 0 && got != 0u) chosen = __shfl_sync(0xffffffffu, k, __ffs(got) - 1);
    if (pos != 0u) last_pos = __shfl_sync(0xffffffffu, k, 31 - __clz(pos));
    base = __shfl_sync(0xffffffffu, pw, 31);
  }
  if (lane == 0) *x_out_row = (chosen >= 0) ? chosen : (last_pos >= 0 ? last_pos : c.xt);
}

// One row: returns the new id (valid in every thread).  `NoiseT` provides E_k when sampling.
template <int NT, typename T, class Row, class NoiseT>
__device__ __forceinline__ int jump_row_math(Row& row, const JumpParams& p, const JumpRowCtx& c, const NoiseT& nz,
                                             const float z_xt, RedRing& red, T* p_row_out, int* argmax_p) {
  const bool exact = (p.flags & FDDM_JUMP_EXACT) != 0;
  const bool sample = (p.flags & FDDM_JUMP_SAMPLE) != 0;
  const bool write_p = (p.flags & FDDM_JUMP_WRITE_P) != 0;
  // softmax in the logits dtype (F.softmax, sampler:189): exp(z-m)/S, rounded to T.
  // FAST (in-kernel RNG: the drawn ids cannot be compared with the reference's anyway) uses MUFU
  // ex2 and a multiply by 1/S; otherwise libm expf and an IEEE division so that the probabilities
  // that decide argmax / the injected-noise race are the reference's bit for bit wherever possible.
  constexpr bool FAST = std::is_same<NoiseT, NoisePhilox>::value;
  constexpr float kLog2e = 1.4426950408889634f;
  float S, inv_S;
  float s2[1] = {0.0f};
  float pm = -1.0f;
  int pm_k = 0x7fffffff;
  const bool want_amax = (argmax_p != nullptr);
  if (FAST) {
    // per-thread max and exp-sum, ONE block reduction of the (max, sum) pairs, then
    // p_k = exp(z_k - m_t) * (exp(m_t - m) / S)
    float m_t = kNegInf;
    row.for_each([&](int, float& x) { m_t = fmaxf(m_t, x); });
    const float nm_t = -m_t * kLog2e;
    float s_t = 0.0f;
    row.for_each([&](int, float& x) {
      x = ex2_approx(fmaf(x, kLog2e, nm_t));
      s_t += x;
    });
    float m = m_t;
    S = s_t;
    block_softmax_stats<NT>(m, S, red);
    inv_S = rcp_approx(S);
    const float sc = ex2_approx((m_t - m) * kLog2e) * inv_S;
    row.for_each([&](int k, float& x) {
      x = Vec16<T>::round_trip(x * sc);
      if (want_amax && x > pm) { pm = x; pm_k = k; }
    });
    s2[0] = S * inv_S;                               // sum of p (== 1 up to rounding): no reduction needed
  } else {
    float m = kNegInf;
    row.for_each([&](int, float& x) { m = fmaxf(m, x); });
    m = block_max<NT>(m, red);
    float s1[1] = {0.0f};
    row.for_each([&](int, float& x) {
      x = expf(x - m);
      s1[0] += x;
    });
    block_sum<NT, 1>(s1, red);
    S = s1[0];
    inv_S = 1.0f / S;
    row.for_each([&](int k, float& x) {
      x = Vec16<T>::round_trip(__fdiv_rn(x, S));
      s2[0] += x;
      if (want_amax && x > pm) { pm = x; pm_k = k; }
    });
  }
  if (write_p) row.store(p_row_out, [](int, float x) { return x; });
  if (argmax_p != nullptr) {
    block_argmax<NT>(pm, pm_k, red);
    *argmax_p = pm_k;
  }
  if (c.identity) return c.xt;                                     // sched:133-134 (delta <= 0)

  // un-normalised target distribution val_k, evaluated on the fly
  float bs = 0.0f, A_gen = 0.0f, A_xt = 0.0f, ab = 0.0f, mixu = 0.0f;
  float sum_xh = 0.0f;
  if (exact) {
    if (!FAST) block_sum<NT, 1>(s2, red);
    sum_xh = s2[0];
    bs = __fmul_rn(c.b_g, sum_xh);                                 // b_tgt * sum(x0hat)     sched:191
    A_gen = c.b_c;                                                 // a*0 + b*sum_xt*1       sched:187
    A_xt = __fadd_rn(c.a_c, c.b_c);
  } else {
    ab = c.ab;
    const float uT = Vec16<T>::round_trip(p.u);                    // full((1,1,K), 1/K, dtype)  sampler:147
    mixu = Vec16<T>::round_trip(__fmul_rn(__fsub_rn(1.0f, ab), uT));   // (1-abar)*u         sampler:151
  }
  auto value = [&](int k, float xh) -> float {
    if (exact) {
      const float Bk = FAST ? fmaf(c.a_g, xh, bs) : __fadd_rn(__fmul_rn(c.a_g, xh), bs);
      return __fmul_rn(k == c.xt ? A_xt : A_gen, Bk);
    }
    if (FAST && sizeof(T) == 4) return fmaf(ab, xh, mixu);
    return Vec16<T>::round_trip(__fadd_rn(Vec16<T>::round_trip(__fmul_rn(ab, xh)), mixu));
  };

  float best = -1.0f;
  int best_k = 0x7fffffff;
  if (!sample) {
    row.for_each([&](int k, float& xh) {
      const float v = value(k, xh);
      if (v > best) { best = v; best_k = k; }
    });
  } else {
    // Categorical weights w_k: the un-normalised target distribution, or with a temperature
    // (sampler:159-161) softmax(log(clamp(p_norm, 1e-12)) / tau) evaluated as exp(logit - max).
    const bool temp = (p.temperature != 1.0f);
    float scale = 1.0f, dn = 1.0f, m2 = 0.0f;
    const float inv_tau = 1.0f / p.temperature;
    auto logit_of = [&](int k, float xh) -> float {
      float pn = value(k, xh);
      if (exact) pn = __fdiv_rn(__fdiv_rn(pn, dn), scale);
      return logf(fmaxf(pn, 1e-12f)) * inv_tau;
    };
    if (temp) {
      if (exact) {
        // exact mode needs the normalised posterior (sched:200-204):
        // denom = a*x0hat[xt] + b*sum_xh*sum_xt ; x0hat[xt] is held by exactly one thread
        float d1[1] = {0.0f};
        row.for_each([&](int k, float& xh) { if (k == c.xt) d1[0] = xh; });
        block_sum<NT, 1>(d1, red);
        dn = fmaxf(__fadd_rn(__fmul_rn(c.a_c, d1[0]), __fmul_rn(c.b_c, sum_xh)), p.eps);
        float ps[1] = {0.0f};
        row.for_each([&](int k, float& xh) { ps[0] += __fdiv_rn(value(k, xh), dn); });
        block_sum<NT, 1>(ps, red);
        scale = fmaxf(ps[0], p.eps);
      }
      m2 = kNegInf;
      row.for_each([&](int k, float& xh) { m2 = fmaxf(m2, logit_of(k, xh)); });
      m2 = block_max<NT>(m2, red);
    }
    auto weight = [&](int k, float xh) -> float { return temp ? expf(logit_of(k, xh) - m2) : value(k, xh); };
    if constexpr (std::is_same<NoiseT, NoisePhilox>::value) {
      // In-kernel RNG: hierarchical exponential race.  The minimum of independent exponentials with
      // rates w_k is Exp(sum w_k) and its argmin is categorical in w_k, independent of the minimum;
      // so each thread races once with its local mass (one Exp(1) variate per thread instead of
      // one per vocab entry) and the winning thread picks among its own entries by inverse CDF.
      float mass = 0.0f;
      row.for_each([&](int k, float& xh) { mass += weight(k, xh); });
      const uint4 rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(row.tid), nz.row, nz.off.x, nz.off.y), nz.key);
      best = __fdividef(mass, exp1_from_bits(rnd.x));
      best_k = row.tid;
      block_argmax<NT>(best, best_k, red);
      int picked = 0;
      if (row.tid == best_k) {
        const float target = mass * ((static_cast<float>(rnd.y >> 8) + 1.0f) * (1.0f / 16777216.0f));
        float cum = 0.0f;
        int chosen = -1, last_pos = c.xt;
        row.for_each([&](int k, float& xh) {
          const float w = weight(k, xh);
          if (chosen < 0 && w > 0.0f) {
            cum += w;
            last_pos = k;
            if (cum >= target) chosen = k;
          }
        });
        picked = (chosen < 0) ? last_pos : chosen;
      }
      return block_broadcast_int<NT>(row.tid == best_k, picked, red);
    } else {
      // injected noise: the reference's per-entry race argmax_k w_k / E_k, bit for bit
      for_each_noise(row, nz, [&](int k, float& xh, float E) {
        const float sc = __fdiv_rn(weight(k, xh), E);
        if (sc > best) { best = sc; best_k = k; }
      });
    }
  }
  block_argmax<NT>(best, best_k, red);
  return best_k;
}

__device__ __forceinline__ void jump_load_ctx(const JumpParams& p, int row, JumpRowCtx& c) {
  const int b = row / p.L;
  c.xt = static_cast<int>(p.x_t[row]);
  c.r0 = c.r1 = c.r2 = 0u;
  c.identity = false;
  c.a_c = c.b_c = c.a_g = c.b_g = 0.0f;
  c.ab = 1.0f;
  if (p.flags & FDDM_JUMP_EXACT) {
    c.a_c = p.coeffs[b];
    c.b_c = p.coeffs[p.B + b];
    c.a_g = p.coeffs[2 * p.B + b];
    c.b_g = p.coeffs[3 * p.B + b];
    c.identity = p.coeffs[4 * p.B] != 0.0f;
  } else {
    c.ab = (p.abar_index < 0) ? 1.0f : p.alpha_bar[p.abar_index];
  }
}

template <int NT>
__device__ __forceinline__ void jump_epilogue(const JumpParams& p, int tid) {
  consumer_sync<NT>();
  if (tid == 0) {
    __threadfence();
    const unsigned int d = atomicAdd(&p.work[1], 1u);
    if (d == gridDim.x - 1) {
      p.work[0] = 0;
      p.work[1] = 0;
      if (p.list_mode) p.work[2] = 0;            // the list has been consumed
      __threadfence();
    }
  }
}

// fast path: TMA ring (logits row [+ noise row] per stage) + register-resident rows
// CTAS = resident CTAs per SM the kernel is compiled for (register budget): the in-kernel-RNG flavour keeps
// few live registers besides the row itself, so more rows can be in flight per SM
template <typename T, int NT, int EPT, int NOISE /*0 none, 1 memory, 2 philox*/, int CTAS>
__global__ void __launch_bounds__(NT + 32, CTAS)
jump_rows_ring_kernel(const JumpParams p, const int nstages, const uint32_t stage_bytes, const uint32_t noise_off) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ uint64_t s_full[kMaxStages], s_empty[kMaxStages];
  __shared__ RingMeta s_meta[kMaxStages];
  __shared__ float s_red[kRedFloats];
  __shared__ int s_nw[32];               // valid row entries owned by each consumer warp (row independent)

  Ring ring;
  ring.stages = dyn_smem;
  ring.stage_bytes = stage_bytes;
  ring.nstages = nstages;
  ring.full = s_full;
  ring.empty = s_empty;
  ring.meta = s_meta;
  ring_init<NT>(ring);

  const int tid = threadIdx.x;
  const uint32_t row_bytes = static_cast<uint32_t>(p.K) * sizeof(T);
  const uint32_t noise_bytes = static_cast<uint32_t>(p.K) * sizeof(float);

  if (tid >= NT) {
    if (tid == NT) {       // one producer lane is enough: no per-row reduction here
      int s = 0;
      uint32_t round = 0;
      for (;;) {
        if (round > 0) mbar_wait_backoff(&ring.empty[s], (round - 1) & 1);
        int row = static_cast<int>(atomicAdd(&p.work[0], 1u));
        const int bound = p.list_mode ? static_cast<int>(__ldcg(&p.work[2])) : p.rows;
        if (row >= bound) {
          ring.meta[s].row = -1;
          mbar_arrive(&ring.full[s]);
          break;
        }
        if (p.list_mode) row = __ldcg(&p.row_list[row]);
        // the copies go out first; the row's metadata (and its random bits) are prepared under their latency;
        // the stage becomes visible to the consumers only with the arrive at the end
        mbar_expect_tx(&ring.full[s], row_bytes + (NOISE == 1 ? noise_bytes : 0u));
        tma_load_1d(ring.stage(s), static_cast<const uint8_t*>(p.logits) + static_cast<size_t>(row) * row_bytes,
                    row_bytes, &ring.full[s]);
        if (NOISE == 1)
          tma_load_1d(ring.stage(s) + noise_off, p.noise + static_cast<size_t>(row) * p.K, noise_bytes, &ring.full[s]);
        JumpRowCtx c;
        jump_load_ctx(p, row, c);
        RingMeta mt;
        mt.row = row; mt.w = c.identity ? 0.0f : 1.0f; mt.i0 = c.xt; mt.i1 = 0;
        mt.f0 = (p.flags & FDDM_JUMP_EXACT) ? c.a_c : c.ab;
        mt.f1 = c.b_c; mt.f2 = c.a_g; mt.f3 = c.b_g;
        mt.r0 = mt.r1 = mt.r2 = 0u;
        if (NOISE == 2) {
          uint2 key, off;
          philox_key_off(p.philox_state, p.key, p.off, key, off);
          const uint4 rnd = philox4x32_10(make_uint4(0xffffffffu, static_cast<uint32_t>(row), off.x, off.y ^ kJumpDomain), key);
          mt.r0 = rnd.x; mt.r1 = rnd.y; mt.r2 = rnd.z;
        }
        ring.meta[s] = mt;
        mbar_arrive(&ring.full[s]);
        if (++s == nstages) { s = 0; ++round; }
      }
    }
    return;
  }

  RedRing red{s_red, 0};
  RegRow<T, NT, EPT> row;
  row.tid = tid;
  row.nvec = p.K / RegRow<T, NT, EPT>::N;
  {
    int n = row.n_owned();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((tid & 31) == 0) s_nw[tid >> 5] = n;       // read only after the first row's block barrier
  }
  const bool fast = (NOISE == 2) && p.temperature == 1.0f;
  int s = 0;
  uint32_t round = 0;
  for (;;) {
    mbar_wait(&ring.full[s], round & 1);
    const RingMeta mt = ring.meta[s];
    if (mt.row < 0) break;
    JumpRowCtx c;
    c.xt = mt.i0; c.identity = (mt.w == 0.0f);
    c.r0 = mt.r0; c.r1 = mt.r1; c.r2 = mt.r2;
    c.a_c = mt.f0; c.ab = mt.f0; c.b_c = mt.f1; c.a_g = mt.f2; c.b_g = mt.f3;
    row.load_from_smem(ring.stage(s), p.K, tid);
    const float z_xt = Vec16<T>::load1(reinterpret_cast<const T*>(ring.stage(s)) + c.xt);
    if (NOISE != 1) ring_release(ring, s);
    T* p_row = (p.flags & FDDM_JUMP_WRITE_P) ? static_cast<T*>(p.p_out) + static_cast<size_t>(mt.row) * p.K : nullptr;
    int amax = 0;
    int id = 0;
    if (NOISE == 2) {
      NoisePhilox nz;
      philox_key_off(p.philox_state, p.key, p.off, nz.key, nz.off); nz.row = static_cast<uint32_t>(mt.row);
      nz.off.y ^= kJumpDomain;
      if (fast)
        jump_row_fast<NT, T>(row, p, c, z_xt, red, s_nw, p_row, p.x_out + mt.row,
                             p.argmax_p_out ? p.argmax_p_out + mt.row : nullptr);
      else
        id = jump_row_math<NT, T>(row, p, c, nz, z_xt, red, p_row, p.argmax_p_out ? &amax : nullptr);
    } else {
      NoiseMem nz;
      nz.p = reinterpret_cast<const float*>(ring.stage(s) + noise_off);
      nz.vec_ok = true;
      id = jump_row_math<NT, T>(row, p, c, nz, z_xt, red, p_row, p.argmax_p_out ? &amax : nullptr);
    }
    if (NOISE == 1) ring_release(ring, s);
    if (tid == 0 && !fast) {
      p.x_out[mt.row] = id;
      if (p.argmax_p_out) p.argmax_p_out[mt.row] = amax;
    }
    if (++s == nstages) { s = 0; ++round; }
  }
  jump_epilogue<NT>(p, tid);
}

constexpr int kStreamStages = 4;

// ------------------------------------------------------------------------------------------------
// In-kernel-RNG production kernel ("streamed"): the row is never held in registers.  One shared-memory
// stage per CTA; the consumers stream over it twice (max, then exp-sum), release it -- the producer's
// next copy then runs under the reductions and the draw -- and the ONE lane that finally needs individual
// entries (the picked lane of the picked warp) re-reads its <= 16 vectors from global memory, where the
// row was streamed microseconds ago (an L2 hit).  ~40 registers per thread instead of ~100, so the number
// of resident CTAs -- rows in flight per SM -- is set by shared memory (6 at V=8000 fp32) and not by the
// register file (4): per-row latency (barrier, reductions, draw) is what bounded the register-resident
// kernel, measured 0.52-0.67 of HBM peak at 3-4 CTAs.  One block barrier per row; hierarchical
// inverse-CDF draw as in jump_row_fast.
// ------------------------------------------------------------------------------------------------
template <typename T, int NT, int CTAS>
__global__ void __launch_bounds__(NT + 32, CTAS) jump_rows_streamed_kernel(const JumpParams p) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ uint64_t s_full[kStreamStages], s_empty[kStreamStages];
  __shared__ RingMeta s_meta[kStreamStages];
  __shared__ float s_red[kRedFloats];
  __shared__ int s_nw[32];
  constexpr int N = Vec16<T>::N, NW = NT / 32;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr float k2m24 = 1.0f / 16777216.0f;
  const int tid = threadIdx.x;
  const int nstages = p.nstages, nchunks = p.nchunks;
  const uint32_t chunk_bytes = static_cast<uint32_t>(p.chunk_vecs) * 16u;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NW);
    }
    mbar_fence_init();
    fence_proxy_async();
  }
  __syncthreads();
  const uint32_t row_bytes = static_cast<uint32_t>(p.K) * sizeof(T);
  const bool exact = (p.flags & FDDM_JUMP_EXACT) != 0;

  if (tid >= NT) {
    if (tid == NT) {                       // producer lane
      uint32_t it = 0;                     // chunks issued so far: stage = it % nstages
      for (;;) {
        const int row = static_cast<int>(atomicAdd(&p.work[0], 1u));
        if (row >= p.rows) {
          const int s = static_cast<int>(it % nstages);
          if (it >= static_cast<uint32_t>(nstages)) mbar_wait_backoff(&s_empty[s], (it / nstages - 1) & 1);
          s_meta[s].row = -1;
          mbar_arrive(&s_full[s]);
          break;
        }
        const uint8_t* src = static_cast<const uint8_t*>(p.logits) + static_cast<size_t>(row) * row_bytes;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = static_cast<int>(it % nstages);
          if (it >= static_cast<uint32_t>(nstages)) mbar_wait_backoff(&s_empty[s], (it / nstages - 1) & 1);
          const uint32_t off_b = static_cast<uint32_t>(c) * chunk_bytes;
          const uint32_t bytes = min(chunk_bytes, row_bytes - off_b);
          mbar_expect_tx(&s_full[s], bytes);
          tma_load_1d(dyn_smem + static_cast<size_t>(s) * chunk_bytes, src + off_b, bytes, &s_full[s]);
          if (c == 0) {                    // the row's metadata travels with its first chunk
            JumpRowCtx cx;
            jump_load_ctx(p, row, cx);
            RingMeta mt;
            mt.row = row; mt.w = cx.identity ? 0.0f : 1.0f; mt.i0 = cx.xt; mt.i1 = 0;
            mt.f0 = exact ? cx.a_c : cx.ab;
            mt.f1 = cx.b_c; mt.f2 = cx.a_g; mt.f3 = cx.b_g;
            uint2 key, off;
            philox_key_off(p.philox_state, p.key, p.off, key, off);
            const uint4 rnd = philox4x32_10(make_uint4(0xffffffffu, static_cast<uint32_t>(row), off.x, off.y ^ kJumpDomain), key);
            mt.r0 = rnd.x; mt.r1 = rnd.y; mt.r2 = rnd.z;
            s_meta[s] = mt;
          }
          mbar_arrive(&s_full[s]);
        }
      }
    }
    return;
  }

  const int lane = tid & 31, warp = tid >> 5;
  const int nvec = p.K / N;
  const int n_own_vec = (nvec - tid + NT - 1) / NT > 0 ? (nvec - tid + NT - 1) / NT : 0;
  {
    int n = n_own_vec * N;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) s_nw[warp] = n;         // read only after the first row's block barrier
  }
  RedRing red{s_red, 0};
  const bool debug_w = (p.flags & FDDM_JUMP_DEBUG_W) != 0;      // test hook: holds the stage like need_p
  const bool need_p = ((p.flags & FDDM_JUMP_WRITE_P) != 0 || p.argmax_p_out != nullptr) && !debug_w;
  // a one-chunk row stays in its stage for the extra pass that writes p_x0 (or the test hook's weights);
  // multi-chunk rows hand every chunk back at once and that pass re-reads the row from global memory (L2)
  const bool hold = (need_p || debug_w) && nchunks == 1;

  uint32_t it = 0;                         // chunks consumed so far
  for (;;) {
    int s0 = static_cast<int>(it % nstages);
    mbar_wait(&s_full[s0], (it / nstages) & 1);
    const RingMeta mt = s_meta[s0];
    if (mt.row < 0) break;
    const int xt = mt.i0;
    const uint4* grow = reinterpret_cast<const uint4*>(static_cast<const T*>(p.logits) + static_cast<size_t>(mt.row) * p.K);
    const uint4* sv = reinterpret_cast<const uint4*>(dyn_smem + static_cast<size_t>(s0) * chunk_bytes);
    float m_t = kNegInf, s_t = 0.0f, z_xt = 0.0f;
    for (int c = 0; c < nchunks; ++c, ++it) {
      const int s = static_cast<int>(it % nstages);
      if (c > 0) mbar_wait(&s_full[s], (it / nstages) & 1);
      const uint4* cv = reinterpret_cast<const uint4*>(dyn_smem + static_cast<size_t>(s) * chunk_bytes);
      const int v0 = c * p.chunk_vecs;
      const int nv_c = min(p.chunk_vecs, nvec - v0);
      // pass 1: thread max over the chunk (four independent chains)
      float mx[4] = {kNegInf, kNegInf, kNegInf, kNegInf};
#pragma unroll 4
      for (int vi = tid; vi < nv_c; vi += NT) {
        float f[N];
        Vec16<T>::unpack(cv[vi], f);
#pragma unroll
        for (int e = 0; e < N; ++e) mx[e & 3] = fmaxf(mx[e & 3], f[e]);
      }
      const float m_new = fmaxf(m_t, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
      // pass 2: thread sum of exp(z - m_new); the running sum is rescaled to the new thread max
      const float nm_t = -m_new * kLog2e;
      float sx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
      for (int vi = tid; vi < nv_c; vi += NT) {
        float f[N];
        Vec16<T>::unpack(cv[vi], f);
#pragma unroll
        for (int e = 0; e < N; ++e) sx[e & 3] += ex2_approx(fmaf(f[e], kLog2e, nm_t));
      }
      s_t = fmaf(s_t, ex2_approx((m_t - m_new) * kLog2e), (sx[0] + sx[1]) + (sx[2] + sx[3]));
      m_t = m_new;
      if (xt >= v0 * N && xt < (v0 + nv_c) * N) z_xt = Vec16<T>::load1(reinterpret_cast<const T*>(cv) + (xt - v0 * N));
      if (!hold) {                         // hand the stage back: the next copy starts now
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
      }
    }
    const uint4* psrc = hold ? sv : grow;  // where the optional third pass reads the row
    // warp (max, sum), one block barrier, row (m, S) and the warp masses in every thread
    float m_w = m_t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, o));
    const float f_tw = ex2_approx((m_t - m_w) * kLog2e);
    float s_w = s_t * f_tw;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s_w += __shfl_xor_sync(0xffffffffu, s_w, o);
    float* sc = red.next();
    if (lane == 0) { sc[warp] = m_w; sc[32 + warp] = s_w; }
    consumer_sync<NT>();
    float m = sc[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = fmaxf(m, sc[w]);
    float sw[NW];
    float S = 0.0f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      sw[w] = sc[32 + w] * ex2_approx((sc[w] - m) * kLog2e);
      S += sw[w];
    }
    const float inv_S = rcp_approx(S);
    const float nm = -m * kLog2e;

    if (need_p) {                          // last jump of a chain: p_x0 in the logits dtype (Q11) and its argmax
      T* p_row = (p.flags & FDDM_JUMP_WRITE_P) ? static_cast<T*>(p.p_out) + static_cast<size_t>(mt.row) * p.K : nullptr;
      float pm = -1.0f;
      int pm_k = 0x7fffffff;
#pragma unroll 4
      for (int vi = tid; vi < nvec; vi += NT) {
        float f[N];
        Vec16<T>::unpack(psrc[vi], f);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          f[e] = Vec16<T>::round_trip(ex2_approx(fmaf(f[e], kLog2e, nm)) * inv_S);
          if (f[e] > pm) { pm = f[e]; pm_k = vi * N + e; }
        }
        if (p_row) stg_stream_v4(reinterpret_cast<uint4*>(p_row) + vi, Vec16<T>::pack(f));
      }
      if (hold) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s0]);
      }
      if (p.argmax_p_out) {
        block_argmax<NT>(pm, pm_k, red);
        if (tid == 0) p.argmax_p_out[mt.row] = pm_k;
      }
    }
    if (mt.w == 0.0f) {                    // sched:133-134 (delta <= 0): identity
      if (tid == 0) p.x_out[mt.row] = xt;
      if (debug_w) {                       // the target distribution is the one-hot of x_t
        T* w_row = static_cast<T*>(p.p_out) + static_cast<size_t>(mt.row) * p.K;
        for (int k = tid; k < p.K; k += NT) Vec16<T>::store1(w_row + k, k == xt ? 1.0f : 0.0f);
        if (hold) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_empty[s0]);
        }
      }
      continue;
    }

    // w_k = wa * p_k + wb, except (wa_x, wb_x) at k == x_t
    const float p_xt = ex2_approx(fmaf(z_xt, kLog2e, nm)) * inv_S;
    float wa, wb, wa_x, wb_x;
    if (exact) {
      const float a_c = mt.f0, b_c = mt.f1, a_g = mt.f2, b_g = mt.f3;
      wa = b_c * a_g; wb = b_c * b_g;                             // A = b_cum * sum_xt, B = a_tgt p + b_tgt * 1
      wa_x = (a_c + b_c) * a_g; wb_x = (a_c + b_c) * b_g;         // A = a_cum + b_cum at x_t   sched:187-191
    } else {
      wa = wa_x = mt.f0;
      wb = wb_x = (1.0f - mt.f0) * p.u;                           // sampler:147-151
    }
    const float corr = fmaf(wa_x - wa, p_xt, wb_x - wb);
    const int owner = (xt / N) % NT;
    float tot = 0.0f;
    float cw[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      float mw = fmaf(wa, sw[w] * inv_S, wb * static_cast<float>(s_nw[w]));
      if (w == (owner >> 5)) mw += corr;
      tot += fmaxf(mw, 0.0f);
      cw[w] = tot;
    }
    if (debug_w) {                         // test hook: the normalised target distribution, entry by entry
      T* w_row = static_cast<T*>(p.p_out) + static_cast<size_t>(mt.row) * p.K;
      const float inv_tot = 1.0f / tot;
      for (int vi = tid; vi < nvec; vi += NT) {
        float f[N];
        Vec16<T>::unpack(psrc[vi], f);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const float pk = ex2_approx(fmaf(f[e], kLog2e, nm)) * inv_S;
          f[e] = fmaxf((vi * N + e == xt) ? fmaf(wa_x, pk, wb_x) : fmaf(wa, pk, wb), 0.0f) * inv_tot;
        }
        stg_stream_v4(reinterpret_cast<uint4*>(w_row) + vi, Vec16<T>::pack(f));
      }
      if (hold) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s0]);
      }
    }
    const float t1 = tot * ((static_cast<float>(mt.r0 >> 8) + 0.5f) * k2m24);
    int wsel = NW - 1;
#pragma unroll
    for (int w = NW - 1; w >= 0; --w)
      if (cw[w] >= t1 && (w == 0 ? cw[0] > 0.0f : cw[w] > cw[w - 1])) wsel = w;
    if (warp != wsel) continue;

    // the picked warp: lane masses -> lane
    float mass = fmaf(wa, s_t * f_tw * ex2_approx((m_w - m) * kLog2e) * inv_S, wb * static_cast<float>(n_own_vec * N));
    if (tid == owner) mass += corr;
    mass = fmaxf(mass, 0.0f);
    float pre = mass;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += up;
    }
    const float wtot = __shfl_sync(0xffffffffu, pre, 31);
    const float t2 = wtot * ((static_cast<float>(mt.r1 >> 8) + 0.5f) * k2m24);
    unsigned hit = __ballot_sync(0xffffffffu, pre >= t2 && mass > 0.0f);
    if (hit == 0u) hit = __ballot_sync(0xffffffffu, mass > 0.0f);
    const int lsel = hit ? (__ffs(hit) - 1) : 0;
    const int tid_sel = (warp << 5) + lsel;
    const float lane_mass = __shfl_sync(0xffffffffu, mass, lsel);
    const float t3 = lane_mass * ((static_cast<float>(mt.r2 >> 8) + 0.5f) * k2m24);
    const int nv_sel = __shfl_sync(0xffffffffu, n_own_vec, lsel);

    // the picked lane's entries: its vectors are re-read from global memory, one per lane
    float base = 0.0f;
    int chosen = -1, last_pos = -1;
    for (int j0 = 0; j0 < nv_sel; j0 += 32) {
      const int j = j0 + lane;
      const bool valid = j < nv_sel;
      const int vi = j * NT + tid_sel;
      float w[N];
      float lsum = 0.0f;
      if (valid) {
        float f[N];
        Vec16<T>::unpack(__ldg(grow + vi), f);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const float pk = ex2_approx(fmaf(f[e], kLog2e, nm)) * inv_S;
          w[e] = fmaxf((vi * N + e == xt) ? fmaf(wa_x, pk, wb_x) : fmaf(wa, pk, wb), 0.0f);
          lsum += w[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < N; ++e) w[e] = 0.0f;
      }
      float pw = lsum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, pw, o);
        if (lane >= o) pw += up;
      }
      pw += base;
      // this lane's own pick, should the target fall into its vector
      const float tl = t3 - (pw - lsum);
      float cum = 0.0f;
      int pick = -1, lastk = -1;
#pragma unroll
      for (int e = 0; e < N; ++e) {
        if (w[e] > 0.0f) {
          cum += w[e];
          lastk = vi * N + e;
          if (pick < 0 && cum >= tl) pick = lastk;
        }
      }
      if (pick < 0) pick = lastk;
      const unsigned pos = __ballot_sync(0xffffffffu, lsum > 0.0f);
      const unsigned got = __ballot_sync(0xffffffffu, lsum > 0.0f && pw >= t3);
      if (chosen < 0 && got != 0u) chosen = __shfl_sync(0xffffffffu, pick, __ffs(got) - 1);
      if (pos != 0u) last_pos = __shfl_sync(0xffffffffu, lastk, 31 - __clz(pos));
      base = __shfl_sync(0xffffffffu, pw, 31);
    }
    if (lane == 0) p.x_out[mt.row] = (chosen >= 0) ? chosen : (last_pos >= 0 ? last_pos : xt);
  }
  jump_epilogue<NT>(p, tid);
}

// ------------------------------------------------------------------------------------------------
// Greedy / posterior_mode == "max" production kernel ("streamed", the reference's default inference
// configuration: configs/fddm_zhTW_base.yaml greedy: true).  argmax_k of the target distribution decides the id,
// and the target is monotone in p_k for every k != x_t, so the decision needs: the row's largest logit, whether
// it is CLEARLY the largest (no other logit within eps_z of it), and -- exact mode -- the one comparison with
// the entry x_t, whose coefficient differs.  All of that is exact-arithmetic-free: a row is decided here only
// when its margin exceeds, by an order of magnitude, every approximation made (MUFU exp / reciprocal, the
// summation order of S, sum(p) = 1) AND the reference's own rounding; then the id equals the reference's id
// bit for bit.  Rows that are not clear (near-ties, degenerate coefficients) are appended to a list and
// re-done by the bit-faithful register-resident kernel (libm expf, IEEE division, the reference's op order),
// which also rewrites their p_x0 / argmax outputs.  Same shape as jump_rows_streamed_kernel: the row stays in
// the stage, two passes, one block barrier, 6 CTAs/SM.
// ------------------------------------------------------------------------------------------------
template <typename T> struct GreedyMargins;
template <> struct GreedyMargins<float> { static constexpr float eps_z = 2.0e-3f, rel = 2.0e-5f; };
template <> struct GreedyMargins<__nv_bfloat16> { static constexpr float eps_z = 4.0e-2f, rel = 1.6e-2f; };
template <> struct GreedyMargins<__half> { static constexpr float eps_z = 6.0e-3f, rel = 2.0e-3f; };

template <typename T, int NT, int CTAS>
__global__ void __launch_bounds__(NT + 32, CTAS) jump_rows_greedy_streamed_kernel(const JumpParams p) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ uint64_t s_full[kStreamStages], s_empty[kStreamStages];
  __shared__ RingMeta s_meta[kStreamStages];
  __shared__ float s_red[kRedFloats];
  constexpr int N = Vec16<T>::N, NW = NT / 32;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr float kEpsZ = GreedyMargins<T>::eps_z, kRel = GreedyMargins<T>::rel;
  const int tid = threadIdx.x;
  const int nstages = p.nstages, nchunks = p.nchunks;
  const uint32_t chunk_bytes = static_cast<uint32_t>(p.chunk_vecs) * 16u;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NW);
    }
    mbar_fence_init();
    fence_proxy_async();
  }
  __syncthreads();
  const uint32_t row_bytes = static_cast<uint32_t>(p.K) * sizeof(T);
  const bool exact = (p.flags & FDDM_JUMP_EXACT) != 0;

  if (tid >= NT) {
    if (tid == NT) {                       // producer lane
      // fast mode at a noisy step: when the uniform mix swamps alpha-bar * p even for p = 1, no row can be
      // clear (see the decision below); every row then goes straight to the bit-faithful kernel's list and
      // nothing is loaded here
      bool hopeless = false;
      if (!exact) {
        const float ab = (p.abar_index < 0) ? 1.0f : p.alpha_bar[p.abar_index];
        hopeless = !(ab > 0.0f) || (ab * kEpsZ * 0.5f < kRel * (ab + (1.0f - ab) * p.u));
      }
      uint32_t it = 0;                     // chunks issued so far: stage = it % nstages
      for (;;) {
        const int row = static_cast<int>(atomicAdd(&p.work[0], 1u));
        if (row >= p.rows) {
          const int s = static_cast<int>(it % nstages);
          if (it >= static_cast<uint32_t>(nstages)) mbar_wait_backoff(&s_empty[s], (it / nstages - 1) & 1);
          s_meta[s].row = -1;
          mbar_arrive(&s_full[s]);
          break;
        }
        if (hopeless) {
          p.row_list[atomicAdd(&p.work[2], 1u)] = row;
          continue;
        }
        const uint8_t* src = static_cast<const uint8_t*>(p.logits) + static_cast<size_t>(row) * row_bytes;
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int s = static_cast<int>(it % nstages);
          if (it >= static_cast<uint32_t>(nstages)) mbar_wait_backoff(&s_empty[s], (it / nstages - 1) & 1);
          const uint32_t off_b = static_cast<uint32_t>(c) * chunk_bytes;
          const uint32_t bytes = min(chunk_bytes, row_bytes - off_b);
          mbar_expect_tx(&s_full[s], bytes);
          tma_load_1d(dyn_smem + static_cast<size_t>(s) * chunk_bytes, src + off_b, bytes, &s_full[s]);
          if (c == 0) {
            JumpRowCtx cx;
            jump_load_ctx(p, row, cx);
            RingMeta mt;
            mt.row = row; mt.w = cx.identity ? 0.0f : 1.0f; mt.i0 = cx.xt; mt.i1 = 0;
            mt.f0 = exact ? cx.a_c : cx.ab;
            mt.f1 = cx.b_c; mt.f2 = cx.a_g; mt.f3 = cx.b_g;
            mt.r0 = mt.r1 = mt.r2 = 0u;
            s_meta[s] = mt;
          }
          mbar_arrive(&s_full[s]);
        }
      }
    }
    return;
  }

  const int lane = tid & 31, warp = tid >> 5;
  const int nvec = p.K / N;
  RedRing red{s_red, 0};
  const bool need_p = (p.flags & FDDM_JUMP_WRITE_P) != 0 || p.argmax_p_out != nullptr;
  const bool hold = need_p && nchunks == 1;               // see jump_rows_streamed_kernel
  const float c_thr = ex2_approx(-kEpsZ * kLog2e);        // e_k >= c_thr  <=>  z_k >= m_t - eps_z

  uint32_t it = 0;                                         // chunks consumed so far
  for (;;) {
    const int s0 = static_cast<int>(it % nstages);
    mbar_wait(&s_full[s0], (it / nstages) & 1);
    const RingMeta mt = s_meta[s0];
    if (mt.row < 0) break;
    const int xt = mt.i0;
    const uint4* grow = reinterpret_cast<const uint4*>(static_cast<const T*>(p.logits) + static_cast<size_t>(mt.row) * p.K);
    const uint4* sv = reinterpret_cast<const uint4*>(dyn_smem + static_cast<size_t>(s0) * chunk_bytes);
    float m_t = kNegInf, s_t = 0.0f, z_xt = 0.0f;
    int cnt_t = 0;
    for (int c = 0; c < nchunks; ++c, ++it) {
      const int s = static_cast<int>(it % nstages);
      if (c > 0) mbar_wait(&s_full[s], (it / nstages) & 1);
      const uint4* cv = reinterpret_cast<const uint4*>(dyn_smem + static_cast<size_t>(s) * chunk_bytes);
      const int v0 = c * p.chunk_vecs;
      const int nv_c = min(p.chunk_vecs, nvec - v0);
      // pass 1: thread max over the chunk
      float mx[4] = {kNegInf, kNegInf, kNegInf, kNegInf};
#pragma unroll 4
      for (int vi = tid; vi < nv_c; vi += NT) {
        float f[N];
        Vec16<T>::unpack(cv[vi], f);
#pragma unroll
        for (int e = 0; e < N; ++e) mx[e & 3] = fmaxf(mx[e & 3], f[e]);
      }
      const float m_new = fmaxf(m_t, fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])));
      // entries of earlier chunks counted as contenders stay contenders only if the thread max did not move
      // beyond eps_z (otherwise they are all below the new threshold); keeping them when in doubt is conservative
      if (m_new > m_t + kEpsZ) cnt_t = 0;
      // pass 2: thread sum of exp(z - m_new) and the number of entries within eps_z of the thread max
      const float nm_t = -m_new * kLog2e;
      float sx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
      for (int vi = tid; vi < nv_c; vi += NT) {
        float f[N];
        Vec16<T>::unpack(cv[vi], f);
#pragma unroll
        for (int e = 0; e < N; ++e) {
          const float ek = ex2_approx(fmaf(f[e], kLog2e, nm_t));
          sx[e & 3] += ek;
          cnt_t += (ek >= c_thr) ? 1 : 0;
        }
      }
      s_t = fmaf(s_t, ex2_approx((m_t - m_new) * kLog2e), (sx[0] + sx[1]) + (sx[2] + sx[3]));
      m_t = m_new;
      if (xt >= v0 * N && xt < (v0 + nv_c) * N) z_xt = Vec16<T>::load1(reinterpret_cast<const T*>(cv) + (xt - v0 * N));
      if (!hold) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
      }
    }
    const uint4* psrc = hold ? sv : grow;
    // warp: max, rescaled sum, contenders (conservative: counted against the thread maxima), owner lane of the max
    float m_w = m_t;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m_w = fmaxf(m_w, __shfl_xor_sync(0xffffffffu, m_w, o));
    float s_w = s_t * ex2_approx((m_t - m_w) * kLog2e);
    int cnt_w = (m_t >= m_w - kEpsZ) ? cnt_t : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s_w += __shfl_xor_sync(0xffffffffu, s_w, o);
      cnt_w += __shfl_xor_sync(0xffffffffu, cnt_w, o);
    }
    const unsigned own = __ballot_sync(0xffffffffu, m_t == m_w);
    float* sc = red.next();
    if (lane == 0) {
      sc[warp] = m_w; sc[32 + warp] = s_w;
      sc[64 + warp] = __int_as_float(cnt_w); sc[96 + warp] = __int_as_float(__ffs(own) - 1);
    }
    consumer_sync<NT>();
    float m = sc[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) m = fmaxf(m, sc[w]);
    float S = 0.0f;
    int contenders = 0, wmax = -1;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      S += sc[32 + w] * ex2_approx((sc[w] - m) * kLog2e);
      if (sc[w] >= m - kEpsZ) contenders += __float_as_int(sc[64 + w]);
      if (wmax < 0 && sc[w] == m) wmax = w;
    }
    const int lmax = __float_as_int(sc[96 + wmax]);
    const float inv_S = rcp_approx(S);
    const float nm = -m * kLog2e;

    if (need_p) {                          // p_x0 in the logits dtype (Q11); its argmax is the clear top entry
      T* p_row = (p.flags & FDDM_JUMP_WRITE_P) ? static_cast<T*>(p.p_out) + static_cast<size_t>(mt.row) * p.K : nullptr;
      if (p_row) {
#pragma unroll 2
        for (int vi = tid; vi < nvec; vi += NT) {
          float f[N];
          Vec16<T>::unpack(psrc[vi], f);
#pragma unroll
          for (int e = 0; e < N; ++e) f[e] = Vec16<T>::round_trip(ex2_approx(fmaf(f[e], kLog2e, nm)) * inv_S);
          stg_stream_v4(reinterpret_cast<uint4*>(p_row) + vi, Vec16<T>::pack(f));
        }
      }
      if (hold) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s0]);
      }
    }

    // ---- the decision (every thread evaluates it; the owner warp acts on it) ----
    const float p1 = inv_S;                                   // the largest probability: exp(0) / S
    const float pxt = ex2_approx(fmaf(z_xt, kLog2e, nm)) * inv_S;
    bool clear = contenders == 1;
    bool pick_xt = false;
    if (mt.w == 0.0f) {                                       // sched:133-134 (delta <= 0): identity
      pick_xt = true;
      clear = clear || !need_p;                               // the id is x_t regardless; argmax p still needs a clear top
    } else if (exact) {
      const float a_c = mt.f0, b_c = mt.f1, a_g = mt.f2, b_g = mt.f3;
      const float A_gen = b_c, A_xt = a_c + b_c;
      const float v1 = A_gen * fmaf(a_g, p1, b_g), vx = A_xt * fmaf(a_g, pxt, b_g);
      const bool top_is_xt = (z_xt == m);
      // a probability margin of eps_z must translate into a value margin of kRel: (a_g p1) / (a_g p1 + b_g) * eps_z >= kRel
      clear = clear && b_c > 0.0f && a_g > 0.0f && (a_g * p1 * kEpsZ * 0.5f >= kRel * fmaf(a_g, p1, b_g));
      if (top_is_xt) {
        pick_xt = true;
        clear = clear && (a_c >= 8.0f * kRel * A_xt);
      } else {
        pick_xt = vx > v1;
        clear = clear && (fabsf(vx - v1) > 8.0f * kRel * fmaxf(vx, v1));
      }
    } else {
      const float ab = mt.f0, mixu = (1.0f - ab) * p.u;
      clear = clear && ab > 0.0f && (ab * p1 * kEpsZ * 0.5f >= kRel * fmaf(ab, p1, mixu));
    }
    if (!clear) {
      if (tid == 0) p.row_list[atomicAdd(&p.work[2], 1u)] = mt.row;   // re-done by the bit-faithful kernel
      continue;
    }
    if (warp != wmax) continue;
    // the owner warp finds the index of the row's (unique, clear) top entry among the owner lane's vectors,
    // re-read from global memory (an L2 hit), one vector per lane
    const int tid_top = (wmax << 5) + lmax;
    const int nv_top = (nvec - tid_top + NT - 1) / NT;
    int k_top = -1;
    for (int j0 = 0; j0 < nv_top && k_top < 0; j0 += 32) {
      const int j = j0 + lane;
      int found = -1;
      if (j < nv_top) {
        const int vi = j * NT + tid_top;
        float f[N];
        Vec16<T>::unpack(__ldg(grow + vi), f);
#pragma unroll
        for (int e = N - 1; e >= 0; --e)
          if (f[e] == m) found = vi * N + e;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, found >= 0);
      if (hit != 0u) k_top = __shfl_sync(0xffffffffu, found, __ffs(hit) - 1);
    }
    if (lane == 0) {
      if (k_top < 0) {                                        // cannot happen; stay safe
        p.row_list[atomicAdd(&p.work[2], 1u)] = mt.row;
      } else {
        p.x_out[mt.row] = pick_xt ? xt : k_top;
        if (p.argmax_p_out) p.argmax_p_out[mt.row] = k_top;
      }
    }
  }
  jump_epilogue<NT>(p, tid);
}

// generic path
template <typename T, int NT, int NOISE>
__global__ void __launch_bounds__(NT, 1) jump_rows_generic_kernel(const JumpParams p) {
  extern __shared__ __align__(128) uint8_t dyn_smem[];
  __shared__ float s_red[kRedFloats];
  RedRing red{s_red, 0};
  float* srow = reinterpret_cast<float*>(dyn_smem);
  const int tid = threadIdx.x;
  SmemRow<T, NT> row;
  for (int r = blockIdx.x; r < p.rows; r += gridDim.x) {
    JumpRowCtx c;
    jump_load_ctx(p, r, c);
    row.load_from_gmem(srow, static_cast<const T*>(p.logits) + static_cast<size_t>(r) * p.K, p.K, tid);
    const float z_xt = srow[c.xt];
    consumer_sync<NT>();                 // everyone has read z_xt before the row is overwritten in place
    T* p_row = (p.flags & FDDM_JUMP_WRITE_P) ? static_cast<T*>(p.p_out) + static_cast<size_t>(r) * p.K : nullptr;
    int amax = 0;
    int id;
    if (NOISE == 2) {
      NoisePhilox nz;
      philox_key_off(p.philox_state, p.key, p.off, nz.key, nz.off); nz.row = static_cast<uint32_t>(r);
      nz.off.y ^= kJumpDomain;
      id = jump_row_math<NT, T>(row, p, c, nz, z_xt, red, p_row, p.argmax_p_out ? &amax : nullptr);
    } else {
      NoiseMem nz;
      nz.p = (NOISE == 1) ? p.noise + static_cast<size_t>(r) * p.K : nullptr;
      nz.vec_ok = false;
      id = jump_row_math<NT, T>(row, p, c, nz, z_xt, red, p_row, p.argmax_p_out ? &amax : nullptr);
    }
    if (tid == 0) {
      p.x_out[r] = id;
      if (p.argmax_p_out) p.argmax_p_out[r] = amax;
    }
    consumer_sync<NT>();
  }
}

template <typename T, int NOISE>
int launch_jump(const JumpParams& p, cudaStream_t stream) {
  const size_t row_bytes = static_cast<size_t>(p.K) * sizeof(T);
  const size_t noise_bytes = (NOISE == 1) ? static_cast<size_t>(p.K) * sizeof(float) : 0;
  bool aligned = (row_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(p.logits) % 16 == 0) &&
                 (!(p.flags & (FDDM_JUMP_WRITE_P | FDDM_JUMP_DEBUG_W)) || reinterpret_cast<uintptr_t>(p.p_out) % 16 == 0);
  if (NOISE == 1) aligned = aligned && (noise_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(p.noise) % 16 == 0);
  const int sms = row_kernel_sms();
  // (the greedy scope spans the streamed kernel AND the bit-faithful pass over its fallback list, which also
  //  records itself under its own name)
  KernelScope ks(p.list_mode ? "jump_rows_bitfaithful_list"
                             : (NOISE == 0 ? "jump_rows_greedy" : (NOISE == 1 ? "jump_rows_injected_noise" : "jump_rows_philox")),
                 stream);
  // Streamed kernels: rows up to 32 KB are one chunk in one stage (6 CTAs/SM at 32 KB); longer rows are copied in
  // 16 KB chunks through two stages with running (max, sum) per thread -- the same 32 KB of shared memory per CTA,
  // so six rows are in flight per SM whatever the vocabulary size (and no vocabulary limit on this path).
  constexpr int kNTs = 128;
  const int nvec_row = p.K / Vec16<T>::N;
  JumpParams ps = p;
  if (row_bytes <= 32 * 1024) { ps.chunk_vecs = (nvec_row + kNTs - 1) / kNTs * kNTs; ps.nchunks = 1; ps.nstages = 1; }
  else { ps.chunk_vecs = 1024; ps.nchunks = (nvec_row + 1023) / 1024; ps.nstages = 2; }
  const size_t stream_smem = static_cast<size_t>(ps.chunk_vecs) * 16 * ps.nstages + 128;
  int stream_ctas = std::min<int>(6, static_cast<int>((216 * 1024) / (stream_smem + 2048)));
  const int stream_grid = static_cast<int>(std::min<int64_t>(p.rows, static_cast<int64_t>(sms) * stream_ctas));
  // the pass that writes p_x0 re-reads a multi-chunk row from global memory; measured at V=32000 that is slower
  // than the register-resident kernel (0.52 vs 0.67 of HBM peak), so the last jump of a chain over long rows
  // stays on the register-resident kernel while it fits (K <= 32768)
  const bool wants_p = (p.flags & (FDDM_JUMP_WRITE_P | FDDM_JUMP_DEBUG_W)) != 0 || p.argmax_p_out != nullptr;
  const bool stream_ok = stream_ctas >= 2 && !(wants_p && ps.nchunks > 1 && p.K <= 32768);
  if (NOISE == 2 && aligned && p.temperature == 1.0f && p.work != nullptr && stream_ok) {
#define FDDM_JUMP_STREAMED(CTAS_)                                                                           \
  do {                                                                                                      \
    auto kfn = jump_rows_streamed_kernel<T, kNTs, CTAS_>;                                                   \
    FDDM_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(stream_smem))); \
    kfn<<<stream_grid, kNTs + 32, stream_smem, stream>>>(ps);                                               \
  } while (0)
    if (stream_ctas >= 6) FDDM_JUMP_STREAMED(6);
    else if (stream_ctas >= 4) FDDM_JUMP_STREAMED(4);
    else FDDM_JUMP_STREAMED(2);
#undef FDDM_JUMP_STREAMED
    FDDM_LAUNCH_OK();
    return FDDM_OK;
  }
  if (NOISE == 0 && aligned && p.K <= 32768 && p.work != nullptr && !p.list_mode && p.row_list != nullptr &&
      stream_ok) {
    // greedy streamed kernel; the rows it does not decide go through the bit-faithful kernel below (list mode)
#define FDDM_JUMP_GREEDY(CTAS_)                                                                             \
  do {                                                                                                      \
    auto kfn = jump_rows_greedy_streamed_kernel<T, kNTs, CTAS_>;                                            \
    FDDM_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(stream_smem))); \
    kfn<<<stream_grid, kNTs + 32, stream_smem, stream>>>(ps);                                               \
  } while (0)
    if (stream_ctas >= 6) FDDM_JUMP_GREEDY(6);
    else if (stream_ctas >= 4) FDDM_JUMP_GREEDY(4);
    else FDDM_JUMP_GREEDY(2);
#undef FDDM_JUMP_GREEDY
    FDDM_LAUNCH_OK();
    JumpParams q = p;
    q.list_mode = 1;
    return launch_jump<T, NOISE>(q, stream);
  }
  if (aligned && p.K <= 32768 && p.work != nullptr) {
    // (consumer threads, row entries per thread, resident CTAs per SM).  The per-row fixed cost (reductions,
    // barrier, draw) is per THREAD, so the in-kernel-RNG flavour uses few threads with many entries each.
    int nt, ept, ctas;
    if (p.K <= 4096) { nt = 128; ept = 32; ctas = (NOISE == 2) ? 4 : 2; }
    else if (p.K <= 8192) { nt = (NOISE == 2) ? 128 : 256; ept = (NOISE == 2) ? 64 : 32; ctas = (NOISE == 2) ? 3 : 2; }
    else if (p.K <= 16384) { nt = 512; ept = 32; ctas = 1; }
    else { nt = 512; ept = 64; ctas = 1; }
    const size_t row_pad = (row_bytes + 127) & ~size_t(127);
    const RingPlan plan = plan_ring(row_pad + noise_bytes, nt, ctas, 2);
    if (plan.nstages >= 1) {
      // (list mode: the number of listed rows is only known on the device -- it can be all of them, e.g. fast
      //  mode at a noisy step where the uniform mix swamps every difference -- so the grid stays full; CTAs
      //  that find the list empty leave at once)
      const int grid = static_cast<int>(std::min<int64_t>(p.rows, static_cast<int64_t>(sms) * plan.ctas_per_sm));
      const uint32_t sb = static_cast<uint32_t>((row_pad + noise_bytes + 127) & ~size_t(127));
#define FDDM_JUMP_LAUNCH(NT_, EPT_, CTAS_)                                                                  \
  do {                                                                                                      \
    auto kfn = jump_rows_ring_kernel<T, NT_, EPT_, NOISE, CTAS_>;                                           \
    FDDM_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,                     \
                                      static_cast<int>(plan.smem_bytes)));                                  \
    kfn<<<grid, NT_ + 32, plan.smem_bytes, stream>>>(p, plan.nstages, sb, static_cast<uint32_t>(row_pad));  \
  } while (0)
      if (NOISE == 2 && nt == 128 && ept == 32) FDDM_JUMP_LAUNCH(128, 32, 4);
      else if (NOISE == 2 && nt == 128 && ept == 64) FDDM_JUMP_LAUNCH(128, 64, 3);
      else if (nt == 128) FDDM_JUMP_LAUNCH(128, 32, 2);
      else if (nt == 256) FDDM_JUMP_LAUNCH(256, 32, 2);
      else if (ept == 32) FDDM_JUMP_LAUNCH(512, 32, 1);
      else FDDM_JUMP_LAUNCH(512, 64, 1);
#undef FDDM_JUMP_LAUNCH
      FDDM_LAUNCH_OK();
      return FDDM_OK;
    }
  }
  const size_t smem = static_cast<size_t>(p.K) * sizeof(float) + 128;
  auto kfn = jump_rows_generic_kernel<T, 256, NOISE>;
  FDDM_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const int grid = static_cast<int>(std::min<int64_t>(p.rows, static_cast<int64_t>(sms) * 4));
  kfn<<<grid, 256, smem, stream>>>(p);
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

}  // namespace

#if FDDM_JUMP_DT == 0
#define FDDM_JUMP_T float
#define FDDM_JUMP_FN jump_launch_f32
#elif FDDM_JUMP_DT == 1
#define FDDM_JUMP_T __nv_bfloat16
#define FDDM_JUMP_FN jump_launch_bf16
#else
#define FDDM_JUMP_T __half
#define FDDM_JUMP_FN jump_launch_f16
#endif
int FDDM_JUMP_FN(const JumpParams& p, int noise, cudaStream_t stream) {
  if (noise == 0) return launch_jump<FDDM_JUMP_T, 0>(p, stream);
  if (noise == 1) return launch_jump<FDDM_JUMP_T, 1>(p, stream);
  return launch_jump<FDDM_JUMP_T, 2>(p, stream);
}
#endif  // FDDM_JUMP_DT
}  // namespace fddm

#ifndef FDDM_JUMP_DT
extern "C" {

size_t fddm_jump_workspace_bytes(int64_t B, int64_t L) {
  if (B <= 0 || L <= 0) return 0;
  return 128 + static_cast<size_t>(B) * static_cast<size_t>(L) * sizeof(int);   // counters + fallback row list
}

int fddm_sample_q_ids(const int64_t* x0, const int64_t* t, const float* alpha_bar, int64_t T, int64_t B, int64_t L,
                      int64_t K, float eps, const float* exp_noise, uint64_t seed, uint64_t offset,
                      const uint64_t* philox_state, int64_t* xt_out, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(x0 && t && alpha_bar && xt_out, "sample_q_ids: null pointer argument");
  FDDM_CHECK_ARG(B > 0 && L > 0 && K > 1 && T > 0, "sample_q_ids: non-positive size");
  FDDM_CHECK_ARG(B * L < (1ll << 31) && K < (1ll << 30), "sample_q_ids: size too large");
  SampleQParams p;
  p.x0 = x0; p.t = t; p.alpha_bar = alpha_bar; p.noise = exp_noise; p.out = xt_out;
  p.T = static_cast<int>(T); p.L = static_cast<int>(L); p.K = static_cast<int>(K);
  p.rows = static_cast<int>(B * L);
  p.eps = eps;
  p.u = static_cast<float>(1.0 / static_cast<double>(K));
  p.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  p.off = make_uint2(static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32));
  p.philox_state = philox_state;
  KernelScope ks(exp_noise ? "sample_q_kernel" : "sample_q_closed_kernel", stream);
  if (exp_noise) {
    const int grid = static_cast<int>(std::min<int64_t>(p.rows, static_cast<int64_t>(num_sms()) * 8));
    sample_q_kernel<256, false><<<grid, 256, 0, stream>>>(p);
  } else {
    sample_q_closed_kernel<<<(p.rows + 255) / 256, 256, 0, stream>>>(p);
  }
  FDDM_LAUNCH_OK();
  return FDDM_OK;
}

int fddm_jump_step(const void* logits, int dtype, const int64_t* x_t, const float* coeffs, const float* alpha_bar,
                   int64_t abar_index, int64_t B, int64_t L, int64_t K, int flags, float temperature, float eps,
                   const float* exp_noise, uint64_t seed, uint64_t offset, const uint64_t* philox_state,
                   void* workspace, int64_t* x_out, int64_t* argmax_p_out, void* p_x0_out, fddm_stream_t stream_) {
  FDDM_API_RANGE();
  using namespace fddm;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FDDM_CHECK_ARG(logits && x_t && x_out, "jump_step: null pointer argument");
  FDDM_CHECK_ARG(dtype_valid(dtype), "jump_step: bad dtype %d", dtype);
  FDDM_CHECK_ARG(B > 0 && L > 0 && K > 1, "jump_step: non-positive size");
  FDDM_CHECK_ARG(B * L < (1ll << 31), "jump_step: too many rows");
  FDDM_CHECK_ARG(!(flags & FDDM_JUMP_EXACT) || coeffs, "jump_step: exact mode needs coeffs");
  FDDM_CHECK_ARG((flags & FDDM_JUMP_EXACT) || alpha_bar || abar_index < 0, "jump_step: fast mode needs alpha_bar");
  FDDM_CHECK_ARG(!(flags & (FDDM_JUMP_WRITE_P | FDDM_JUMP_DEBUG_W)) || p_x0_out, "jump_step: WRITE_P needs p_x0_out");
  FDDM_CHECK_ARG(!(flags & FDDM_JUMP_DEBUG_W) || ((flags & FDDM_JUMP_SAMPLE) && !exp_noise && temperature == 1.0f),
                 "jump_step: DEBUG_W is a hook of the in-kernel-RNG flavour (SAMPLE, no injected noise, temperature 1)");
  FDDM_CHECK_ARG(temperature > 0.0f, "jump_step: temperature must be positive");
  if (K > FDDM_MAX_VOCAB) {
    set_error("jump_step: K=%lld exceeds FDDM_MAX_VOCAB=%d", (long long)K, FDDM_MAX_VOCAB);
    return FDDM_EUNSUPPORTED;
  }
  JumpParams p;
  p.logits = logits; p.x_t = x_t; p.coeffs = coeffs; p.alpha_bar = alpha_bar; p.noise = exp_noise;
  p.x_out = x_out; p.argmax_p_out = argmax_p_out; p.p_out = p_x0_out;
  p.work = static_cast<unsigned int*>(workspace);
  p.row_list = workspace ? reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + 128) : nullptr;
  p.list_mode = 0;
  p.chunk_vecs = 0; p.nchunks = 1; p.nstages = 1;
  p.B = static_cast<int>(B); p.L = static_cast<int>(L); p.K = static_cast<int>(K); p.rows = static_cast<int>(B * L);
  p.flags = flags; p.abar_index = static_cast<int>(abar_index);
  p.temperature = temperature; p.eps = eps;
  p.u = static_cast<float>(1.0 / static_cast<double>(K));
  p.key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  p.off = make_uint2(static_cast<uint32_t>(offset), static_cast<uint32_t>(offset >> 32));
  p.philox_state = philox_state;
  const bool sample = (flags & FDDM_JUMP_SAMPLE) != 0;
  const int noise = !sample ? 0 : (exp_noise ? 1 : 2);
  if (dtype == FDDM_F32) return jump_launch_f32(p, noise, stream);
  if (dtype == FDDM_BF16) return jump_launch_bf16(p, noise, stream);
  return jump_launch_f16(p, noise, stream);
}

}  // extern "C"
#endif  // !FDDM_JUMP_DT
