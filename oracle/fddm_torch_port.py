"""CPU baseline port of the reference's token path in PyTorch-eager form.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/fddm_oracle.py's header): only `tests/` and `bench.py`'s
baseline legs (`cpu_baseline`, `--impl reference`, and `eager_b200` = this same file run with device="cuda"
as the "PyTorch-eager on B200" bar of SURVEY.md section 2b) import this file; nothing under `fddm-asr_b200/` does.
Every function works on whatever device its inputs live on.

The reference (TeemoCaption/FDDM-asr) is pure PyTorch and cannot travel to the GPU box, so the CPU arm
of the benchmark times this restatement instead (`cpu_baseline.kind = "port"`).  It performs the same
sequence of dense [B,L,V] torch operations the reference performs -- one-hot scatter, softmax, the
closed-form transition products, un-normalised posteriors, logs, reductions, autograd backward,
torch.multinomial / Categorical -- so that its cost on the host cores is the reference's cost, using
all intra-op threads torch is given.  Parity pinning: tests/test_oracle_golden.py checks it against
the same reference-generated golden vectors as the numpy oracle.

Citations: sched = fddm/sched/diffusion_scheduler.py, train = train.py, losses = losses/fddm_losses.py,
sampler = sampler/jumpy_sampler.py of the reference.
"""
from __future__ import annotations

import math

import torch


def make_tables(K: int, T: int, beta_max: float = 0.2, device="cpu"):
    """sched:25-29."""
    steps = torch.arange(1, T + 1, device=device, dtype=torch.float32)
    betas = beta_max * torch.sin(0.5 * math.pi * (steps / float(T))) ** 2
    return betas, torch.cumprod(1.0 - betas, dim=0)


@torch.no_grad()
def q_sample(x0_prob, t, alpha_bar, eps=1e-8):
    """sched:44-49."""
    B, L, K = x0_prob.shape
    ab = alpha_bar[t - 1].view(B, 1, 1)
    xt = ab * x0_prob.float() + (1.0 - ab) * torch.full_like(x0_prob, 1.0 / K, dtype=torch.float32)
    xt = xt.clamp_min(eps)
    return xt / xt.sum(dim=-1, keepdim=True).clamp_min(eps)


@torch.no_grad()
def sample_q(x0, t, alpha_bar, K):
    """train:180-188: ids -> one-hot -> q_sample -> multinomial."""
    B, L = x0.shape
    oh = torch.zeros(B, L, K, device=x0.device)
    oh.scatter_(-1, x0.unsqueeze(-1), 1.0)
    p = q_sample(oh, t, alpha_bar)
    return torch.multinomial(p.view(-1, K), 1).view(B, L)


def kl_term(xt, x0, logits, t, betas, x_mask=None):
    """train:190-255 (differentiable w.r.t. logits)."""
    B, L, V = logits.shape
    dtype = logits.dtype
    xh = torch.softmax(logits, dim=-1)
    xt_oh = torch.zeros(B, L, V, dtype=dtype, device=logits.device).scatter_(-1, xt.unsqueeze(-1), 1.0)
    x0_oh = torch.zeros(B, L, V, dtype=dtype, device=logits.device).scatter_(-1, x0.unsqueeze(-1), 1.0)
    beta_t = betas[t - 1]
    beta_p = torch.where(t.eq(1), torch.zeros_like(beta_t), betas[(t - 2).clamp(min=0)])
    bt, bp = beta_t.view(B, 1, 1), beta_p.view(B, 1, 1)
    one = torch.ones_like(xh)
    eps, Kf = 1e-8, float(V)
    u = (bt / Kf) * one + (1.0 - bt) * xt_oh
    v = (1.0 - bp) * x0_oh + (bp / Kf) * one
    w = (1.0 - bp) * xh + (bp / Kf) * one
    same = torch.sum(x0_oh * xt_oh, dim=-1)
    xh_at = torch.gather(xh, -1, xt.unsqueeze(-1)).squeeze(-1)
    b1 = beta_t.unsqueeze(-1)
    d_q = (b1 / Kf) + (1.0 - b1) * same
    d_p = (b1 / Kf) + (1.0 - b1) * xh_at
    q = (u * v) / (d_q.unsqueeze(-1) + eps)
    p = (u * w) / (d_p.unsqueeze(-1) + eps)
    kl_tok = torch.sum(q * (torch.log(q + eps) - torch.log(p + eps)), dim=-1)
    if x_mask is not None:
        valid = x_mask.float()
        per = (kl_tok * valid).sum(dim=1) / (valid.sum(dim=1) + eps)
    else:
        per = kl_tok.mean(dim=1)
    return per.mean()


@torch.no_grad()
def multistep_coeffs(t, delta, betas, K, T):
    """sched:132-183 as executed (aliased in-place recurrence, quirk Q1), scalar Python loop like
    the reference's."""
    B = t.shape[0]
    dev = t.device
    a_c, b_c = torch.ones(B, device=dev), torch.zeros(B, device=dev)
    a_g, b_g = torch.ones(B, device=dev), torch.zeros(B, device=dev)
    tl = t.tolist()
    bl = betas.tolist()
    f32 = lambda x: torch.tensor(x, dtype=torch.float32)
    for i in range(B):
        a = torch.tensor(1.0); b = torch.tensor(0.0)
        tt = max(tl[i] - delta, 0)
        for s in range(tl[i], tt, -1):
            if 1 <= s <= T:
                a_s = 1.0 - f32(bl[s - 1]); b_s = f32(bl[s - 1]) / K
                a = a_s * a
                b = a_s * b + b_s * (a + K * b)
        a_c[i], b_c[i] = a, b
        if 0 < tt <= T:
            a_g[i] = 1.0 - f32(bl[tt - 1]); b_g[i] = f32(bl[tt - 1]) / K
    return a_c, b_c, a_g, b_g


@torch.no_grad()
def q_posterior_multi_step(xt_prob, xh, t, delta, betas, T, eps=1e-8):
    """sched:106-208."""
    B, L, K = xt_prob.shape
    delta = min(delta, int(t.min().item()))
    if delta <= 0:
        return xt_prob
    a_c, b_c, a_g, b_g = [c.view(B, 1, 1) for c in multistep_coeffs(t, delta, betas, K, T)]
    ones = torch.ones_like(xt_prob)
    sum_xt = xt_prob.sum(dim=-1, keepdim=True)
    A = a_c * xt_prob + b_c * sum_xt * ones
    sum_xh = xh.sum(dim=-1, keepdim=True)
    Bt = a_g * xh + b_g * sum_xh * ones
    dot = torch.sum(xt_prob * xh, dim=-1, keepdim=True)
    denom = a_c * dot + b_c * sum_xh * sum_xt
    post = (A * Bt) / denom.clamp_min(eps)
    return post / post.sum(dim=-1, keepdim=True).clamp_min(eps)


@torch.no_grad()
def jump_once(x_t, logits, t_scalar, delta, betas, alpha_bar, K, T_train, T_infer, sampling_mode="exact",
              greedy=True, temperature=1.0):
    """sampler:167-217 given this step's logits."""
    B, L = x_t.shape
    p_x0 = torch.softmax(logits, dim=-1)
    if sampling_mode == "exact":
        oh = torch.zeros(B, L, K, device=x_t.device).scatter_(-1, x_t.unsqueeze(-1), 1.0)
        t_vec = torch.full((B,), t_scalar, dtype=torch.long, device=x_t.device)
        p = q_posterior_multi_step(oh, p_x0.float(), t_vec, delta, betas, T_train)
    else:
        tgt = max(0, t_scalar - delta)
        if tgt <= 0:
            ab = torch.tensor(1.0)
        else:
            idx = int(round(max(1.0, min(float(T_train), tgt / float(max(1, T_infer)) * T_train))))
            ab = alpha_bar[idx]
        p = ab * p_x0 + (1.0 - ab) * torch.full((1, 1, K), 1.0 / K, dtype=p_x0.dtype, device=p_x0.device)
    if greedy:
        return p.argmax(dim=-1), p_x0
    if temperature != 1.0:
        p = torch.softmax(p.clamp(min=1e-12).log() / temperature, dim=-1)
    return torch.distributions.Categorical(probs=p).sample(), p_x0


def lfd_loss(z_a, z_b, lambda_offdiag=5e-3, eps=1e-5):
    """losses:18-58."""
    B, T, D = z_a.shape

    def std(x):
        mean = x.mean(dim=0, keepdim=True)
        var = x.var(dim=0, unbiased=False, keepdim=True)
        return (x - mean) / torch.sqrt(var + eps)

    za, zb = std(z_a).reshape(B * T, D), std(z_b).reshape(B * T, D)
    C = (za.T @ zb) / (B * T)
    diag = torch.diagonal(C)
    off = C - torch.diag(diag)
    return torch.sum((1.0 - diag) ** 2) + lambda_offdiag * torch.sum(off ** 2)
