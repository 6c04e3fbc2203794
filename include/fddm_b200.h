/*
 * fddm_b200.h -- C ABI of libfddm_b200.so: the B200-native (sm_100a) implementation of FDDM-ASR's
 * categorical discrete-diffusion token path.
 *
 * Every entry point takes raw DEVICE pointers, int64 sizes and a cudaStream_t; no torch types, no
 * C++ exceptions, no allocation and no host synchronisation inside the library.  All buffers are
 * caller-owned and only borrowed for the duration of the stream-ordered call; inputs are never
 * mutated.  Return value: FDDM_OK (0) or a negative fddm_status_t; fddm_last_error() gives the
 * thread-local message of the last failure.  All calls are CUDA-graph capturable.
 *
 * Reference interface replaced by each function is cited as <file>:<line> in
 * TeemoCaption/FDDM-asr (sched = fddm/sched/diffusion_scheduler.py, train = train.py,
 * losses = losses/fddm_losses.py, sampler = sampler/jumpy_sampler.py).
 *
 * Layout conventions: logits / probability tensors are row-major contiguous [B, L, V] ("rows" =
 * B*L token rows of V vocab entries); ids are int64 [B, L]; t is int64 [B] with values in 1..T;
 * betas / alpha_bar are fp32 [T], 0-based (alpha_bar[t-1] is the paper's alpha-bar_t, sched:25-29).
 */
#ifndef FDDM_B200_H_
#define FDDM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef FDDM_NO_CUDA_RUNTIME_TYPES
typedef struct CUstream_st* fddm_stream_t;      /* == cudaStream_t */
#endif

typedef enum { FDDM_F32 = 0, FDDM_BF16 = 1, FDDM_F16 = 2 } fddm_dtype_t;

typedef enum {
  FDDM_OK = 0,
  FDDM_EINVAL = -1,        /* bad argument (null pointer, non-positive size, bad enum, misalignment) */
  FDDM_ECUDA = -2,         /* a CUDA runtime call / launch failed; message holds cudaGetErrorString   */
  FDDM_EUNSUPPORTED = -3   /* shape outside the supported envelope (e.g. V > FDDM_MAX_VOCAB)          */
} fddm_status_t;

#define FDDM_MAX_VOCAB 49152       /* one fp32 row must fit the 227 KB shared memory of an SM */
#define FDDM_ABI_VERSION 3

/* jump_step flags */
#define FDDM_JUMP_EXACT   0x1      /* sampling_mode == "exact" (else "fast"), sampler:192-209 */
#define FDDM_JUMP_SAMPLE  0x2      /* Categorical sampling (else argmax), sampler:153-162 / 212-215 */
#define FDDM_JUMP_WRITE_P 0x4      /* also write softmax(logits) = p_x0 in the logits dtype, sampler:189 */
#define FDDM_JUMP_DEBUG_W 0x8      /* test hook of the in-kernel-RNG flavour (SAMPLE, no injected noise, temperature 1):
                                      p_x0_out receives, instead of p_x0, the NORMALISED target distribution the draw
                                      is taken from (the Delta-step posterior / alpha-bar mix), so that the fast
                                      arithmetic of that flavour can be compared with the oracle entry by entry */

int fddm_version(void);
const char* fddm_last_error(void);
/* number of kernels launched by this library in this process since load (bench evidence) */
int64_t fddm_launch_count(void);

/* Measurement aid (bench.py's per-kernel roofline block).  While enabled, every kernel launch of the
 * library is bracketed by a pair of CUDA events on its stream (never during stream capture); every entry
 * point and launch is also an NVTX range, always.  fddm_profile_enable(1) clears earlier records.
 * fddm_profile_read synchronises the recorded events and writes "kernel<TAB>launches<TAB>total_ms" lines
 * into buf (NUL-terminated) when cap suffices; returns the bytes needed. */
int fddm_profile_enable(int on);

/* The persistent row kernels (KL, jump step) normally fill every SM.  A caller that runs a collective
 * (NCCL all-reduce) concurrently on another stream reserves the SMs that collective's CTAs need, so that it
 * is never queued behind the persistent CTAs: the row kernels then launch (SMs - n) * resident CTAs and their
 * dynamic row scheduler does the rest.  Process-wide; 0 restores the default. */
int fddm_set_sm_reserve(int n);
int64_t fddm_profile_read(char* buf, int64_t cap);

/* ------------------------------------------------------------------------------------------------
 * a2  DiscreteDiffusionScheduler.q_sample                                     sched:31-50
 *   out[b,l,:] = normalise(clamp_min(abar_t * x0_prob + (1-abar_t)/K, eps))
 * x0_prob, out: fp32 [B,L,K].  12 B/element would be 2 reads; this is 1 read + 1 write = 8 B. */
int fddm_q_sample_dense(const float* x0_prob, const int64_t* t, const float* alpha_bar, int64_t T,
                        int64_t B, int64_t L, int64_t K, float eps, float* out, fddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a3  SchedulerAdapter.sample_q  (ids -> ids, the one-hot is never materialised)  train:180-188
 *   xt = argmax_k q_sample(onehot(x0), t)_k / E_k         ( == torch.multinomial(p, 1) )
 * exp_noise: fp32 [B*L, K] of Exp(1) variates (the per-entry race replays torch.multinomial bit for
 * bit), or NULL to draw in-kernel with Philox4x32-10 keyed by (seed, offset): the row of q_sample is
 * two-valued for a one-hot x0, so the draw is O(1) per token (keep x0 with probability p_hi, else a
 * uniform other id).  philox_state: optional device uint64[2] = {seed, offset} that overrides the two
 * scalars (so a captured CUDA graph draws fresh noise on every replay).  Ties -> lowest index. */
int fddm_sample_q_ids(const int64_t* x0, const int64_t* t, const float* alpha_bar, int64_t T,
                      int64_t B, int64_t L, int64_t K, float eps, const float* exp_noise,
                      uint64_t seed, uint64_t offset, const uint64_t* philox_state, int64_t* xt_out,
                      fddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a4  DiscreteDiffusionScheduler.q_posterior                                  sched:52-104
 * xt_prob, x0hat_prob, out: fp32 [B,L,K]. */
int fddm_q_posterior_dense(const float* xt_prob, const float* x0hat_prob, const int64_t* t,
                           const float* betas, int64_t T, int64_t B, int64_t L, int64_t K, float eps,
                           float* out, fddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a5  q_posterior_multi_step, part 1: the per-sample transition-product coefficients
 *                                                                            sched:132-183
 * Replaces the reference's Python double loop and its 2B+1 .item() host syncs with one tiny
 * kernel.  delta_eff = min(delta, min_b t[b], *delta_cap if given) is computed on the device
 * (the batch-global clip of sched:132; delta_cap lets a batch-sharded caller pass the global
 * min).  coeffs: fp32 [4*B + 1] = a_cum[B] | b_cum[B] | a_tgt[B] | b_tgt[B] | identity flag
 * (1.0 when delta_eff <= 0, in which case consumers return xt unchanged, sched:133-134). */
int fddm_multistep_coeffs(const int64_t* t, int64_t B, int64_t delta, const int64_t* delta_cap,
                          const float* betas, int64_t T, int64_t K, float* coeffs,
                          fddm_stream_t stream);

/* a5 part 2: the dense posterior given the coefficients                        sched:186-206 */
int fddm_q_posterior_multi_dense(const float* xt_prob, const float* x0hat_prob, const float* coeffs,
                                 int64_t B, int64_t L, int64_t K, float eps, float* out,
                                 fddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a7  SchedulerAdapter.kl_term                                                train:190-255
 * One fused kernel per token row: vocab softmax, both transition products, the two un-normalised
 * posteriors (+eps, quirk Q5), KL, the masked per-sample mean and the batch mean; the backward
 * variant additionally writes d loss / d logits in the same pass (logits are read exactly once).
 *
 *   logits      [B,L,V] dtype `dtype`           xt, x0  int64 [B,L]        t  int64 [B]
 *   x_mask      NULL (plain mean over L), or [B,L] uint8/bool (non-zero = valid), or -- flags & FDDM_KL_MASK_F32 --
 *               [B,L] fp32 WEIGHTS: the reference multiplies by x_mask.float() (train:250), so a non-boolean mask
 *               acts as per-token weights sum_l w KL / (sum_l w + eps)
 *   flags       FDDM_KL_MASK_F32 | FDDM_KL_CLAMP_T.  Without CLAMP_T a t outside 1..T makes the kernel print the
 *               offending value and trap (the reference raises: IndexError on CPU, device-side assert on CUDA);
 *               with it t is clamped silently
 *   batch_div   the divisor of the final batch mean (B for one process; the GLOBAL batch when the
 *               batch is sharded over ranks -- then *loss_out is this rank's partial sum / batch_div
 *               and an all-reduce SUM completes it)
 *   workspace   fddm_kl_workspace_bytes(B,L) bytes, zero-initialised ONCE by the caller
 *   loss_out    fp32 scalar
 *   grad_scale  fp32 device scalar multiplied into the gradient (upstream dL/dloss), NULL = 1
 *   grad_logits [B,L,V] dtype `dtype` (the reference's grad has the logits dtype) */
#define FDDM_KL_MASK_F32 0x1
#define FDDM_KL_CLAMP_T  0x2
size_t fddm_kl_workspace_bytes(int64_t B, int64_t L);
int fddm_kl_forward(const void* logits, int dtype, const int64_t* xt, const int64_t* x0,
                    const int64_t* t, const void* x_mask, int flags, const float* betas, int64_t T,
                    int64_t B, int64_t L, int64_t V, double batch_div, void* workspace,
                    float* loss_out, fddm_stream_t stream);
int fddm_kl_forward_backward(const void* logits, int dtype, const int64_t* xt, const int64_t* x0,
                             const int64_t* t, const void* x_mask, int flags, const float* betas, int64_t T,
                             int64_t B, int64_t L, int64_t V, double batch_div,
                             const float* grad_scale, void* workspace, float* loss_out,
                             void* grad_logits, fddm_stream_t stream);
/* x[i] *= (*num) / (den ? *den : 1); every CTA exits before touching x when the ratio is exactly 1
 * (autograd fix-up when the upstream gradient differs from the one assumed by the fused pass). */
int fddm_scale_inplace(void* x, int dtype, int64_t n, const float* num, const float* den,
                       fddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a9  DiffusionJumpySampler._jump_once minus the decoder call                 sampler:189-215
 * One fused kernel per token row: softmax(logits) -> Delta-step posterior given one-hot x_t
 * (exact; coefficients from fddm_multistep_coeffs) or alpha-bar mix with uniform (fast) ->
 * argmax or exponential-race sample (== Categorical.sample()) -> new ids; optionally p_x0.
 *
 *   flags        FDDM_JUMP_* bits
 *   coeffs       exact mode: the [4B+1] array of fddm_multistep_coeffs
 *   alpha_bar    fast mode: table [T]; abar_index = the (quirk-Q3, 0-based) table index to use, or
 *                -1 for alpha-bar = 1 (target step <= 0)                       sampler:219-236
 *   temperature  used only when sampling and != 1                              sampler:159-161
 *   exp_noise    fp32 [B*L, K] Exp(1) variates (per-entry race, bit-faithful arithmetic) or NULL ->
 *                in-kernel Philox (seed, offset) or device {seed, offset} in philox_state: hierarchical
 *                exponential race, one variate per thread, MUFU arithmetic
 *   workspace    fddm_jump_workspace_bytes(B,L) bytes whose first 128 are zero-initialised ONCE by the caller
 *                (self-resetting row scheduler counters; the rest is the list of rows the greedy fast kernel
 *                hands to the bit-faithful kernel)
 *   argmax_p_out optional int64 [B,L]: argmax_k p_x0 -- the sampler's final x_0 (sampler:292), fused
 *                here so the last p_x0 is never re-read
 *   p_x0_out     [B,L,K] in the logits dtype when FDDM_JUMP_WRITE_P, else may be NULL */
size_t fddm_jump_workspace_bytes(int64_t B, int64_t L);
int fddm_jump_step(const void* logits, int dtype, const int64_t* x_t, const float* coeffs,
                   const float* alpha_bar, int64_t abar_index, int64_t B, int64_t L, int64_t K,
                   int flags, float temperature, float eps, const float* exp_noise, uint64_t seed,
                   uint64_t offset, const uint64_t* philox_state, void* workspace, int64_t* x_out,
                   int64_t* argmax_p_out, void* p_x0_out, fddm_stream_t stream);
/* ------------------------------------------------------------------------------------------------
 * a8  lfd_loss                                                                losses:18-58
 * z_a, z_b: [B,T,D] dtype `dtype` (this rank's batch shard); rows = B*T.  Phases are separate entry
 * points so that a batch-sharded caller can all-reduce (SUM) between them; with one process the
 * same sequence runs without collectives:
 *
 *   1. fddm_lfd_stats     sums = fp64 [2 tensors][2 moments][T*D]: sum_b x, sum_b x^2   (losses:23-24)
 *                         -> all-reduce sums
 *   2. fddm_lfd_xcov      cov[D,D] = sum_rows za~^T zb~ over this rank's rows, fp32.  One pass standardises
 *                         za~ = (x-mean)/sqrt(var+eps) with the GLOBAL-batch mean/var (n_batch_global
 *                         samples), splits fp32 into bf16 hi+residual (3 MMAs, ~2^-16) and writes packed
 *                         operand planes; a tcgen05 (UMMA, TMEM accumulator) contraction fed by TMA bulk
 *                         copies does the rest; the diagonal is accumulated in fp64.       losses:25-48
 *                         -> all-reduce cov
 *   3. fddm_lfd_loss      C = cov / n_rows_global; loss = sum_j (1-C_jj)^2 + lambda sum_{j!=k} C_jk^2;
 *                         G[D,D] = dloss/dC                                               losses:51-57
 *   4. fddm_lfd_backward  phase 0: dza~ = zb~ G^T / N, dzb~ = za~ G / N (two tcgen05 contractions into
 *                         the workspace) and bn_sums = fp32 [2 tensors][T][P][D], P = fddm_lfd_bn_parts(B,T,D):
 *                         partial batch sums of dz~ z~ per (t, d) (for B >= 32 they come out of the
 *                         contraction's epilogue, one partial per 32 batch rows; P = 1 for B < 32).
 *                         The other batch-norm moment, sum_b dz~, is identically zero (z~ has zero batch
 *                         mean) and is not computed.
 *                         -> sum the P partials and all-reduce [2][T][D] when the batch is sharded
 *                         phase 1: dz = (dz~ - z~ mean_b(dz~ z~)) / std * (*grad_scale), written in the
 *                         input dtype; bn_parts = P, or 1 when the caller passes already-summed sums.
 *
 * workspace: fddm_lfd_workspace_bytes(B,T,D) bytes whose first 256 bytes are zero-initialised ONCE by
 * the caller (self-resetting counters); the rest is scratch (tables, split-K partials, dz~).
 * D must be a multiple of 8 (FDDM_EUNSUPPORTED otherwise). */
#define FDDM_LFD_PLANES_VALID 0x2  /* OR into phase 0 of fddm_lfd_backward: `workspace` still holds the
                                      operand planes / tables written by THIS call's fddm_lfd_xcov */
size_t fddm_lfd_workspace_bytes(int64_t B, int64_t T, int64_t D);
int fddm_lfd_stats(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D,
                   double* sums, fddm_stream_t stream);
int fddm_lfd_xcov(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D,
                  const double* sums, double n_batch_global, float eps, void* workspace,
                  float* cov /* [D,D] */, fddm_stream_t stream);
int fddm_lfd_loss(const float* cov, int64_t D, double n_rows_global, float lambda_offdiag,
                  void* workspace, float* loss_out, float* G /* [D,D] */, fddm_stream_t stream);
int64_t fddm_lfd_bn_parts(int64_t B, int64_t T, int64_t D);
int fddm_lfd_backward(const void* z_a, const void* z_b, int dtype, int64_t B, int64_t T, int64_t D,
                      const double* sums, double n_batch_global, float eps, const float* G,
                      double n_rows_global, const float* grad_scale, void* workspace,
                      float* bn_sums, int64_t bn_parts, int phase, void* dz_a, void* dz_b,
                      fddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * f4  calculate_cer / calculate_wer                                    models/evaluate.py:94-134
 * Batched Levenshtein distance (unit costs) over integer symbol sequences: characters' code points for CER,
 * word ids for WER.  All pairs of an evaluation set in one launch.
 *   ref, hyp          int32 concatenated symbols            ref_off, hyp_off   int64 [n_pairs + 1] offsets
 *   max_hyp_len       the longest hypothesis (sizes the workspace: fddm_edit_distance_workspace_bytes)
 *   dist_out          int32 [n_pairs] edit distances (the rates are dist / len(ref) with the reference's
 *                     empty-reference rules, applied by the host mirror) */
size_t fddm_edit_distance_workspace_bytes(int64_t n_pairs, int64_t max_hyp_len);
int fddm_edit_distance(const int32_t* ref, const int64_t* ref_off, const int32_t* hyp,
                       const int64_t* hyp_off, int64_t n_pairs, int64_t max_hyp_len, void* workspace,
                       int32_t* dist_out, fddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (e) multi-GPU exchange of the batch-sharded L_fd                     SURVEY.md section 8(e)
 * The library's own SUM all-reduce, in place, of a buffer that lives in symmetric memory (every rank has every
 * peer's copy mapped over NVLink / NVSwitch): one kernel, two cross-GPU flag barriers, no NCCL.  Replaces
 * ncclAllReduce for the three 1.6-6.3 MB exchanges of losses/fddm_losses.py's statistics when the batch is sharded
 * (the reference itself has no multi-GPU path).
 *   FDDM_XGPU_P2P   each rank loads its 1/world slice from every copy (summed in rank order: bitwise the same
 *                   result on every rank) and stores the sums into every copy.  world 2, 4 or 8.
 *   FDDM_XGPU_NVLS  the slice is reduced and re-broadcast by the switch (multimem.ld_reduce / multimem.st)
 *                   through the multicast mapping.
 *   buffer_ptrs_dev      DEVICE array [world] of pointers to the ranks' copies (P2P), 16-byte aligned
 *   multicast_ptr        multicast address of the buffer (NVLS), or NULL
 *   signal_pad_ptrs_dev  DEVICE array [world] of pointers to the ranks' signal pads, each at least
 *                        fddm_xgpu_signal_pad_bytes() long and zero before first use; the library uses bytes
 *                        [4096, fddm_xgpu_signal_pad_bytes()) only
 *   elem_bytes           4 (fp32) or 8 (fp64); n * elem_bytes must be a multiple of 16
 *   max_ctas             upper bound on the CTAs of the kernel (0: the default, 32) -- the number of SMs to
 *                        reserve with fddm_set_sm_reserve when it runs under a persistent row kernel
 * Every rank must make the same call (same n, algo, max_ctas) in the same order; stream-ordered, CUDA-graph
 * capturable.  Two calls that share signal pads must not execute concurrently (order them on one stream or with
 * events, as the host mirror does); a rank that never makes the call leaves the others spinning in the barrier. */
#define FDDM_XGPU_P2P 1
#define FDDM_XGPU_NVLS 2
int64_t fddm_xgpu_signal_pad_bytes(void);
int fddm_xgpu_allreduce(const void* buffer_ptrs_dev, void* multicast_ptr, const void* signal_pad_ptrs_dev, int rank,
                        int world, int elem_bytes, int64_t n, int algo, int max_ctas, fddm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FDDM_B200_H_ */
