"""CPU oracle for the FDDM-ASR categorical discrete-diffusion token path.

TEST INFRASTRUCTURE ONLY.  This file is a numpy restatement of the reference's algorithm
(TeemoCaption/FDDM-asr, pure PyTorch) used as the *checker* for the CUDA path.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import it; nothing under `fddm-asr_b200/` does (the product fails loudly without its
CUDA library, it never falls back to this file).

Parity pinning: the reference ships no golden vectors for this path (its only check,
scripts/sanity_check_scheduler.py:19-26, asserts row sums == 1).  The oracle is therefore
pinned against outputs of the reference itself, produced in the build container by importing
`/root/reference` (script: tests/golden/make_golden.py, fixtures: tests/golden/*.npz), and
`tests/test_oracle_golden.py` checks every function below against them.

The arithmetic of the reference lives in PyTorch (requirements.txt:2, `torch>=2.0.0`; the
fixtures were generated with torch 2.11.0+cu128 on CPU).  Each function mirrors the
reference's fp32 operation order; pass `dtype=np.float64` for the high-precision variant used
as tie-breaker.

Citations (`sched` = fddm/sched/diffusion_scheduler.py, `train` = train.py,
`losses` = losses/fddm_losses.py, `sampler` = sampler/jumpy_sampler.py).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """Round fp32 values to the nearest bf16 (ties to even), returned as fp32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    lsb = (u >> 16) & 1
    r = ((u + 0x7FFF + lsb) >> 16) << 16
    out = r.astype(np.uint32).view(np.float32).copy()
    nan = np.isnan(x)
    if nan.any():
        out[nan] = np.nan
    return out.reshape(x.shape)


def round_to_dtype(x: np.ndarray, io_dtype: str) -> np.ndarray:
    """Round an fp32 array to the value set of `io_dtype` ('f32' | 'bf16' | 'f16')."""
    if io_dtype == "f32":
        return np.asarray(x, dtype=np.float32)
    if io_dtype == "bf16":
        return round_to_bf16(x)
    if io_dtype == "f16":
        return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)
    raise ValueError(io_dtype)


def softmax_lastdim(z: np.ndarray, dtype=F32) -> np.ndarray:
    z = np.asarray(z, dtype=dtype)
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m)
    return e / e.sum(axis=-1, keepdims=True)


def argmax_first(x: np.ndarray) -> np.ndarray:
    """argmax over the last axis, ties -> lowest index (torch.argmax / np.argmax semantics)."""
    return np.argmax(x, axis=-1).astype(np.int64)


# ---------------------------------------------------------------------------------------------
# a1  tables  (sched:18-29)
# ---------------------------------------------------------------------------------------------
def make_tables(K: int, T: int, beta_max: float = 0.2, dtype=F32) -> Tuple[np.ndarray, np.ndarray]:
    """betas[i] = beta_max * sin^2(pi/2 * (i+1)/T), alpha_bar = cumprod(1-betas); 0-based
    (alpha_bar[t-1] is the paper's alpha-bar_t).  sched:25-29."""
    t = np.arange(1, T + 1, dtype=dtype)
    s = np.sin(dtype(0.5 * math.pi) * (t / dtype(T)))
    betas = (dtype(beta_max) * (s * s)).astype(dtype)
    alpha_bar = np.cumprod((dtype(1.0) - betas).astype(dtype), dtype=dtype)
    return betas, alpha_bar


# ---------------------------------------------------------------------------------------------
# a2  q_sample  (sched:31-50)
# ---------------------------------------------------------------------------------------------
def q_sample(x0_prob: np.ndarray, t: np.ndarray, alpha_bar: np.ndarray, eps: float = 1e-8,
             dtype=F32) -> np.ndarray:
    """x_t = abar_t*x0 + (1-abar_t)*(1/K); clamp_min(eps); / max(sum_k, eps).  sched:44-49."""
    x0 = np.asarray(x0_prob, dtype=dtype)
    B, L, K = x0.shape
    ab = np.asarray(alpha_bar, dtype=dtype)[np.asarray(t, dtype=np.int64) - 1].reshape(B, 1, 1)
    u = dtype(1.0 / K)                                  # python double -> fp32 (sched:45)
    xt = ab * x0 + (dtype(1.0) - ab) * u
    xt = np.maximum(xt, dtype(eps))
    s = np.maximum(xt.sum(axis=-1, keepdims=True, dtype=dtype), dtype(eps))
    return (xt / s).astype(dtype)


def q_sample_two_values(t: np.ndarray, alpha_bar: np.ndarray, K: int, eps: float = 1e-8,
                        dtype=F32) -> Tuple[np.ndarray, np.ndarray]:
    """For one-hot x0 the row of q_sample has two distinct values: (p_hi at x0, p_lo elsewhere),
    *before* the final division; also returns nothing else -- callers divide by the row sum.
    Mirrors sched:47-48 on a one-hot input."""
    ab = np.asarray(alpha_bar, dtype=dtype)[np.asarray(t, dtype=np.int64) - 1]
    u = dtype(1.0 / K)
    lo = (dtype(1.0) - ab) * u                           # ab*0 + (1-ab)*u
    hi = ab * dtype(1.0) + lo
    return np.maximum(hi, dtype(eps)), np.maximum(lo, dtype(eps))


# ---------------------------------------------------------------------------------------------
# a3  SchedulerAdapter.sample_q  (train.py:180-188)
# ---------------------------------------------------------------------------------------------
def sample_q_ids(x0: np.ndarray, t: np.ndarray, alpha_bar: np.ndarray, K: int,
                 exp_noise: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """ids -> one-hot -> q_sample -> torch.multinomial(p, 1).
    torch.multinomial(p, 1) == argmax_k(p_k / E_k) with E ~ Exp(1) i.i.d. (ATen
    native/Sampling / MultinomialKernel: q.exponential_(1); p.div_(q); argmax), ties -> lowest
    index.  `exp_noise` is the injected E of shape [B, L, K] (fp32)."""
    x0 = np.asarray(x0, dtype=np.int64)
    B, L = x0.shape
    onehot = np.zeros((B, L, K), dtype=F32)
    np.put_along_axis(onehot, x0[..., None], F32(1.0), axis=-1)
    p = q_sample(onehot, t, alpha_bar, eps)
    return argmax_first(p / np.asarray(exp_noise, dtype=F32).reshape(B, L, K))


# ---------------------------------------------------------------------------------------------
# a4  q_posterior  (sched:52-104)
# ---------------------------------------------------------------------------------------------
def _beta_prev(betas: np.ndarray, t: np.ndarray, dtype=F32) -> np.ndarray:
    """beta_{t-1} with beta_0 := 0  (sched:76-86 / train.py:214-217)."""
    t = np.asarray(t, dtype=np.int64)
    prev = np.asarray(betas, dtype=dtype)[np.clip(t - 2, 0, None)]
    return np.where(t == 1, dtype(0.0), prev).astype(dtype)


def q_posterior(xt_prob: np.ndarray, x0hat_prob: np.ndarray, t: np.ndarray, betas: np.ndarray,
                eps: float = 1e-8, dtype=F32) -> np.ndarray:
    xt = np.asarray(xt_prob, dtype=dtype)
    xh = np.asarray(x0hat_prob, dtype=dtype)
    B, L, K = xt.shape
    t = np.asarray(t, dtype=np.int64)
    beta_t = np.asarray(betas, dtype=dtype)[t - 1].reshape(B, 1, 1)
    a_t = dtype(1.0) - beta_t
    b_t = beta_t / dtype(K)
    beta_p = _beta_prev(betas, t, dtype).reshape(B, 1, 1)
    a_p = dtype(1.0) - beta_p
    b_p = beta_p / dtype(K)
    A = a_t * xt + b_t                                     # sched:91
    Bv = a_p * xh + b_p                                    # sched:94
    dot = (xt * xh).sum(axis=-1, keepdims=True, dtype=dtype)
    denom = a_t * dot + b_t                                # sched:98
    post = (A * Bv) / np.maximum(denom, dtype(eps))        # sched:101
    s = np.maximum(post.sum(axis=-1, keepdims=True, dtype=dtype), dtype(eps))
    return (post / s).astype(dtype)


# ---------------------------------------------------------------------------------------------
# a5  q_posterior_multi_step  (sched:106-208)
# ---------------------------------------------------------------------------------------------
def multistep_coeffs(t: np.ndarray, delta: int, betas: np.ndarray, K: int, T: int,
                     dtype=F32) -> Tuple[int, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Returns (delta_eff, a_cum[B], b_cum[B], a_tgt[B], b_tgt[B]).

    delta_eff = min(delta, t.min())  (sched:132, a *batch-global* min -- quirk Q10).
    The cumulative recurrence is restated AS EXECUTED by the reference (quirk Q1): `a_old` /
    `b_old` at sched:157-158 are 0-dim *views* into the accumulators, so after the in-place
    write of sched:160 the `a_old` read by sched:161 is already the updated value:
        a' = a_s * a ;  b' = a_s * b + b_s * (a' + K * b)
    Target coefficients are the single-step matrix of t_target (sched:168-183), (1, 0) when
    t_target == 0."""
    t = np.asarray(t, dtype=np.int64)
    betas = np.asarray(betas, dtype=dtype)
    B = t.shape[0]
    delta_eff = int(min(int(delta), int(t.min())))
    a_cum = np.ones(B, dtype=dtype)
    b_cum = np.zeros(B, dtype=dtype)
    a_tgt = np.ones(B, dtype=dtype)
    b_tgt = np.zeros(B, dtype=dtype)
    if delta_eff <= 0:
        return delta_eff, a_cum, b_cum, a_tgt, b_tgt
    t_tgt = np.clip(t - delta_eff, 0, None)
    Kf = dtype(K)
    for i in range(B):
        a = dtype(1.0)
        b = dtype(0.0)
        for step in range(int(t[i]), int(t_tgt[i]), -1):
            if 1 <= step <= T:
                beta_s = betas[step - 1]
                a_s = dtype(1.0) - beta_s
                b_s = beta_s / Kf
                a = dtype(a_s * a)                           # sched:160 (in place)
                # sched:161 reads the *updated* a through the aliased view; python int K
                # times an fp32 0-dim tensor stays fp32.
                b = dtype(dtype(a_s * b) + dtype(b_s * dtype(a + dtype(Kf * b))))
        a_cum[i] = a
        b_cum[i] = b
        tt = int(t_tgt[i])
        if 0 < tt <= T:
            a_tgt[i] = dtype(1.0) - betas[tt - 1]
            b_tgt[i] = betas[tt - 1] / Kf
    return delta_eff, a_cum, b_cum, a_tgt, b_tgt


def q_posterior_multi_step(xt_prob: np.ndarray, x0hat_prob: np.ndarray, t: np.ndarray, delta: int,
                           betas: np.ndarray, T: int, eps: float = 1e-8, dtype=F32) -> np.ndarray:
    xt = np.asarray(xt_prob, dtype=dtype)
    xh = np.asarray(x0hat_prob, dtype=dtype)
    B, L, K = xt.shape
    d_eff, a_c, b_c, a_g, b_g = multistep_coeffs(t, delta, betas, K, T, dtype)
    if d_eff <= 0:
        return xt                                           # sched:133-134
    a_c = a_c.reshape(B, 1, 1); b_c = b_c.reshape(B, 1, 1)
    a_g = a_g.reshape(B, 1, 1); b_g = b_g.reshape(B, 1, 1)
    sum_xt = xt.sum(axis=-1, keepdims=True, dtype=dtype)
    A = a_c * xt + b_c * sum_xt                              # sched:187 ( * ones is exact)
    sum_xh = xh.sum(axis=-1, keepdims=True, dtype=dtype)
    Bt = a_g * xh + b_g * sum_xh                             # sched:191
    dot = (xt * xh).sum(axis=-1, keepdims=True, dtype=dtype)
    denom = a_c * dot + b_c * sum_xh * sum_xt                # sched:197
    post = (A * Bt) / np.maximum(denom, dtype(eps))          # sched:200
    s = np.maximum(post.sum(axis=-1, keepdims=True, dtype=dtype), dtype(eps))
    return (post / s).astype(dtype)


# ---------------------------------------------------------------------------------------------
# a6  w_t  (sched:210-213, train.py:257-261)
# ---------------------------------------------------------------------------------------------
def w_t(t: np.ndarray, alpha_bar: np.ndarray) -> np.ndarray:
    return np.asarray(alpha_bar)[np.asarray(t, dtype=np.int64) - 1]


# ---------------------------------------------------------------------------------------------
# a7  SchedulerAdapter.kl_term  (train.py:190-255)
# ---------------------------------------------------------------------------------------------
def kl_token_terms(xt: np.ndarray, x0: np.ndarray, logits: np.ndarray, t: np.ndarray,
                   betas: np.ndarray, io_dtype: str = "f32", dtype=F32, want_grad: bool = False):
    """Per-token KL[q(x_{t-1}|x_t,x_0) || p_theta(x_{t-1}|x_t,c)] of train.py:201-246 (posteriors
    un-normalised, eps=1e-8 added not clamped: quirk Q5) and, if asked, d KL_tok / d logits.

    `io_dtype` is the dtype of `logits` in the reference run; the softmax is evaluated in that
    dtype (its result rounded to it -- quirk Q11, train.py:201) and everything after is fp32
    because the beta tensors are fp32."""
    xt = np.asarray(xt, dtype=np.int64)
    x0 = np.asarray(x0, dtype=np.int64)
    z = np.asarray(logits, dtype=dtype)
    B, L, V = z.shape
    t = np.asarray(t, dtype=np.int64)
    eps = dtype(1e-8)
    Kf = dtype(V)
    xh = softmax_lastdim(z, dtype)
    if dtype is F32 or dtype == np.float32:
        xh = round_to_dtype(xh, io_dtype)
    beta_t = np.asarray(betas, dtype=dtype)[t - 1].reshape(B, 1, 1)
    beta_p = _beta_prev(betas, t, dtype).reshape(B, 1, 1)
    a_t = dtype(1.0) - beta_t
    b_t = beta_t / Kf
    a_p = dtype(1.0) - beta_p
    b_p = beta_p / Kf
    xt_oh = np.zeros((B, L, V), dtype=dtype)
    np.put_along_axis(xt_oh, xt[..., None], dtype(1.0), axis=-1)
    x0_oh = np.zeros((B, L, V), dtype=dtype)
    np.put_along_axis(x0_oh, x0[..., None], dtype(1.0), axis=-1)
    u = b_t + a_t * xt_oh                                    # train.py:227
    v = a_p * x0_oh + b_p                                    # train.py:230
    w = a_p * xh + b_p                                       # train.py:231
    same = (x0 == xt).astype(dtype)[..., None]               # train.py:234
    xh_at_xt = np.take_along_axis(xh, xt[..., None], axis=-1)
    d_q = b_t + a_t * same                                   # train.py:238
    d_p = b_t + a_t * xh_at_xt                               # train.py:239
    q = (u * v) / (d_q + eps)                                # train.py:242
    p = (u * w) / (d_p + eps)                                # train.py:243
    kl_tok = (q * (np.log(q + eps) - np.log(p + eps))).sum(axis=-1, dtype=dtype)   # train.py:246
    if not want_grad:
        return kl_tok, None
    # closed-form gradient wrt logits (verified against reference autograd by make_golden.py)
    P = d_p + eps
    r = q / (p + eps)
    S = (r * p).sum(axis=-1, keepdims=True, dtype=dtype)
    g = -r * u * a_p / P + xt_oh * (a_t * S / P)
    gdot = (g * xh).sum(axis=-1, keepdims=True, dtype=dtype)
    grad = xh * (g - gdot)
    return kl_tok, grad.astype(dtype)


def kl_reduce(kl_tok: np.ndarray, x_mask: Optional[np.ndarray], dtype=F32):
    """Masked per-sample mean then mean over the batch (train.py:249-255, quirk Q6).
    Returns (loss, token_weight[B,L]) with d loss / d kl_tok = token_weight."""
    B, L = kl_tok.shape
    if x_mask is not None:
        valid = np.asarray(x_mask).astype(dtype)
        den = valid.sum(axis=1, dtype=dtype) + dtype(1e-8)
        per = (kl_tok * valid).sum(axis=1, dtype=dtype) / den
        wgt = valid / den[:, None] / dtype(B)
    else:
        per = kl_tok.mean(axis=1, dtype=dtype)
        wgt = np.full((B, L), 1.0 / (L * B), dtype=dtype)
    return dtype(per.mean(dtype=dtype)), wgt.astype(dtype)


def kl_term(xt, x0, logits, t, betas, x_mask=None, io_dtype="f32", dtype=F32, want_grad=False):
    """Returns loss (0-dim fp32) and, if asked, d loss / d logits (rounded to io_dtype)."""
    kl_tok, g = kl_token_terms(xt, x0, logits, t, betas, io_dtype, dtype, want_grad)
    loss, wgt = kl_reduce(kl_tok, x_mask, dtype)
    if not want_grad:
        return loss, None
    grad = g * wgt[..., None]
    if dtype is F32 or dtype == np.float32:
        grad = round_to_dtype(grad, io_dtype)
    return loss, grad


# ---------------------------------------------------------------------------------------------
# a8  lfd_loss  (losses:18-58)
# ---------------------------------------------------------------------------------------------
def lfd_standardize(x: np.ndarray, eps: float = 1e-5, dtype=F32):
    """Standardise along the batch axis per (position, channel).  losses:23-26."""
    x = np.asarray(x, dtype=dtype)
    mean = x.mean(axis=0, keepdims=True, dtype=dtype)
    var = x.var(axis=0, keepdims=True, dtype=dtype)          # biased (unbiased=False)
    std = np.sqrt(var + dtype(eps))
    return ((x - mean) / std).astype(dtype), mean, std


def lfd_loss(z_a: np.ndarray, z_b: np.ndarray, lambda_offdiag: float = 5e-3, eps: float = 1e-5,
             dtype=F32, want_grad: bool = False):
    """C = za~^T zb~ / (B*T);  loss = sum_j (1-C_jj)^2 + lambda * sum_{j!=k} C_jk^2.
    losses:39-57.  Gradient (closed form, verified against reference autograd):
    G_jj = -2(1-C_jj), G_jk = 2*lambda*C_jk; dza~ = zb~ G^T / N, dzb~ = za~ G / N;
    dx = (dz~ - mean_b dz~ - z~ * mean_b(dz~ * z~)) / std."""
    za_in = np.asarray(z_a, dtype=dtype)
    zb_in = np.asarray(z_b, dtype=dtype)
    B, T, D = za_in.shape
    assert zb_in.shape == (B, T, D)
    za, _, sa = lfd_standardize(za_in, eps, dtype)
    zb, _, sb = lfd_standardize(zb_in, eps, dtype)
    N = B * T
    A2 = za.reshape(N, D)
    B2 = zb.reshape(N, D)
    C = (A2.T @ B2) / dtype(N)
    diag = np.diagonal(C)
    off = C - np.diag(diag)
    loss = ((dtype(1.0) - diag) ** 2).sum(dtype=dtype) + dtype(lambda_offdiag) * (off ** 2).sum(dtype=dtype)
    if not want_grad:
        return dtype(loss), None, None
    G = (dtype(2.0 * lambda_offdiag) * off).astype(dtype)
    G[np.arange(D), np.arange(D)] = dtype(-2.0) * (dtype(1.0) - diag)
    dza = (B2 @ G.T / dtype(N)).reshape(B, T, D)
    dzb = (A2 @ G / dtype(N)).reshape(B, T, D)

    def _bn_bwd(dz, zt, std):
        m1 = dz.mean(axis=0, keepdims=True, dtype=dtype)
        m2 = (dz * zt).mean(axis=0, keepdims=True, dtype=dtype)
        return ((dz - m1 - zt * m2) / std).astype(dtype)

    return dtype(loss), _bn_bwd(dza, za, sa), _bn_bwd(dzb, zb, sb)


# ---------------------------------------------------------------------------------------------
# a9 / a10  jumpy sampler  (sampler:139-293)
# ---------------------------------------------------------------------------------------------
def alpha_bar_at_t_train(t_infer: int, T_infer: int, T_train: int, alpha_bar: np.ndarray):
    """sampler:219-236 including the 0-based/1-based off-by-one (quirk Q3): a 1-based
    train-axis index is used directly on the 0-based table (IndexError at index T_train)."""
    if t_infer <= 0:
        return F32(1.0)
    ratio = float(t_infer) / float(max(1, T_infer))
    tf = ratio * float(T_train)
    tf = max(1.0, min(float(T_train), tf))
    idx = int(round(tf))
    return np.asarray(alpha_bar, dtype=F32)[idx]


def categorical_from_probs(p: np.ndarray, exp_noise: np.ndarray) -> np.ndarray:
    """torch.distributions.Categorical(probs=p).sample(): probs are renormalised by their sum
    (distributions/categorical.py, quirk Q12) and sampled with torch.multinomial(.,1) ==
    argmax(p/E)."""
    p = np.asarray(p, dtype=F32)
    p = p / p.sum(axis=-1, keepdims=True, dtype=F32)
    return argmax_first(p / np.asarray(exp_noise, dtype=F32))


def to_indices(p: np.ndarray, greedy: bool, temperature: float, exp_noise: Optional[np.ndarray]):
    """sampler:153-162."""
    if greedy:
        return argmax_first(p)
    if temperature != 1.0:
        lg = np.log(np.maximum(p, F32(1e-12))) / F32(temperature)
        p = softmax_lastdim(lg)
    return categorical_from_probs(p, exp_noise)


def jump_once(x_t_idx: np.ndarray, logits: np.ndarray, t_scalar: int, delta: int, *, K: int,
              T_train: int, T_infer: int, betas: np.ndarray, alpha_bar: np.ndarray,
              sampling_mode: str = "exact", posterior_mode: str = "average", greedy: bool = True,
              temperature: float = 1.0, exp_noise: Optional[np.ndarray] = None,
              io_dtype: str = "f32", eps: float = 1e-8):
    """One `_jump_once` (sampler:167-217) given the decoder's logits for this step.
    exact mode feeds the T_infer-axis `t` to the T_train-length beta table un-rescaled (Q2).
    Returns (x_{t-delta} ids, p_x0, p_xtmd)."""
    x_t_idx = np.asarray(x_t_idx, dtype=np.int64)
    B, L = x_t_idx.shape
    p_x0 = round_to_dtype(softmax_lastdim(np.asarray(logits, dtype=F32)), io_dtype)   # sampler:189
    if sampling_mode == "exact":
        onehot = np.zeros((B, L, K), dtype=F32)
        np.put_along_axis(onehot, x_t_idx[..., None], F32(1.0), axis=-1)
        t_vec = np.full((B,), t_scalar, dtype=np.int64)
        p = q_posterior_multi_step(onehot, p_x0, t_vec, delta, betas, T_train, eps)   # sampler:200
    else:
        ab = alpha_bar_at_t_train(max(0, t_scalar - delta), T_infer, T_train, alpha_bar)
        # sampler:147-151.  With half-precision logits p_x0 is half, `u` is created in
        # p_x0.dtype and the 0-dim fp32 alpha-bar does not promote, so every op of the mix
        # rounds to the half dtype (for 'f32' the roundings are identities).
        u = round_to_dtype(np.array([1.0 / K], dtype=F32), io_dtype)
        p1 = round_to_dtype(F32(ab) * p_x0, io_dtype)
        p2 = round_to_dtype((F32(1.0) - F32(ab)) * u, io_dtype)
        p = round_to_dtype(p1 + p2, io_dtype)
    if posterior_mode == "max":
        ids = argmax_first(p)                                                          # sampler:213
    else:
        ids = to_indices(p, greedy, temperature, exp_noise)                            # sampler:215
    return ids, p_x0, p


def sample_loop(x_T: np.ndarray, logits_fn, *, K: int, T_train: int, T_infer: int, r: int,
                betas, alpha_bar, sampling_mode="exact", posterior_mode="average", greedy=True,
                temperature=1.0, noise_fn=None, io_dtype="f32"):
    """sampler:282-293: jumps of min(r, t); the returned ids are argmax of the LAST p_x0, the last
    jump's resampled ids are discarded (quirk Q9)."""
    x = np.asarray(x_T, dtype=np.int64)
    t = T_infer
    p_last = None
    step = 0
    while t > 0:
        delta = min(r, t)
        noise = noise_fn(step) if (noise_fn is not None) else None
        x, p_last, _ = jump_once(x, logits_fn(x, t, step), t, delta, K=K, T_train=T_train,
                                 T_infer=T_infer, betas=betas, alpha_bar=alpha_bar,
                                 sampling_mode=sampling_mode, posterior_mode=posterior_mode,
                                 greedy=greedy, temperature=temperature, exp_noise=noise,
                                 io_dtype=io_dtype)
        t -= delta
        step += 1
    return argmax_first(p_last), p_last, x


# ---------------------------------------------------------------------------------------------
# near-tie certification (parity tests use it to classify any id mismatch)
# ---------------------------------------------------------------------------------------------
def near_tie(scores_row: np.ndarray, i: int, j: int, ulps: float = 8.0) -> bool:
    """True if candidates i and j of a score row are within `ulps` fp32 ulps of each other."""
    a = float(scores_row[i]); b = float(scores_row[j])
    scale = max(abs(a), abs(b), 1e-45)
    return abs(a - b) <= ulps * scale * 2.0 ** -23


# ---------------------------------------------------------------------------------------------
# CER / WER (models/evaluate.py:94-134)
# ---------------------------------------------------------------------------------------------
def edit_distance(r, h) -> int:
    """Levenshtein distance with unit costs between two sequences (the DP of evaluate.py:101-115)."""
    n, m = len(r), len(h)
    prev = list(range(m + 1))
    for i in range(1, n + 1):
        cur = [i] + [0] * m
        for j in range(1, m + 1):
            cost = 0 if r[i - 1] == h[j - 1] else 1
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + cost)
        prev = cur
    return prev[m]


def calculate_cer(ref: str, hyp: str) -> float:
    """evaluate.py:94-118."""
    if len(ref) == 0:
        return 0.0 if len(hyp) == 0 else 1.0
    return float(edit_distance(list(ref), list(hyp))) / float(len(ref))


def calculate_wer(ref: str, hyp: str) -> float:
    """evaluate.py:120-134."""
    r, h = ref.strip().split(), hyp.strip().split()
    return 0.0 if len(r) == 0 else float(edit_distance(r, h)) / float(len(r))
