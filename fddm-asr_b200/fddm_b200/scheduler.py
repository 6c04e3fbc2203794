"""DiscreteDiffusionScheduler -- B200 host mirror of the reference scheduler.

Same constructor, attributes and method signatures as the reference
`fddm/sched/diffusion_scheduler.py:17-213` (`K, T, device, eps, betas, alpha_bar`, `q_sample`,
`q_posterior`, `q_posterior_multi_step`, `w_prefix`), so `train.py` / `inference.py` /
`sampler/jumpy_sampler.py` of the reference use it unchanged.  Every method is one call into
libfddm_b200.so (include/fddm_b200.h); there is no CPU path.

Extras beyond the reference (keyword-only, optional): `sample_q_ids` (the fused ids->ids forward
corruption the adapter uses) and `multistep_coeffs`.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib as L


class DiscreteDiffusionScheduler:
    def __init__(self, K: int, T: int, device, beta_max: float = 0.2, eps: float = 1e-8):
        self.K = int(K)
        self.T = int(T)
        self.device = device
        self.eps = float(eps)
        # Tables: the same torch expressions as the reference (sched:25-29) so that on the same
        # device they are bit-identical to the reference's.  0-based: alpha_bar[t-1] = abar_t.
        steps = torch.arange(1, T + 1, device=device, dtype=torch.float32)
        self.betas = beta_max * torch.sin(0.5 * math.pi * (steps / float(T))) ** 2
        self.alpha_bar = torch.cumprod(1.0 - self.betas, dim=0)

    # ------------------------------------------------------------------------------------------
    def _prep(self, *probs: torch.Tensor, t: torch.Tensor):
        outs = []
        for p in probs:
            p = p.to(self.device).float()                      # sched:39-40 / 61-63 / 124-126
            if p.dim() != 3:
                raise ValueError(f"expected a (B, L, K) tensor, got shape {tuple(p.shape)}")
            outs.append(p.contiguous())
        t = t.to(self.device).long().contiguous()
        B, Lq, K = outs[0].shape
        assert K == self.K                                     # sched:42 / 65 / 129
        for p in outs[1:]:
            if p.shape != outs[0].shape:
                raise ValueError("probability tensors must have the same shape")
        if t.shape != (B,):
            raise ValueError(f"t must have shape ({B},), got {tuple(t.shape)}")
        L.require_cuda(*outs, t, self.betas)
        return outs, t, B, Lq, K

    @torch.no_grad()
    def q_sample(self, x0_prob: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """q(x_t|x_0) = normalise(clamp_min(abar_t x0 + (1-abar_t)/K, eps)).  sched:31-50."""
        (x0,), t, B, Lq, K = self._prep(x0_prob, t=t)
        out = torch.empty_like(x0)
        L.check(L.lib.fddm_q_sample_dense(x0.data_ptr(), t.data_ptr(), self.alpha_bar.data_ptr(), self.T, B, Lq, K,
                                          self.eps, out.data_ptr(), L.stream_ptr(x0.device)), "q_sample")
        return out

    @torch.no_grad()
    def q_posterior(self, xt_prob: torch.Tensor, x0hat_prob: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """Single-step posterior q(x_{t-1}|x_t, x0hat).  sched:52-104."""
        (xt, xh), t, B, Lq, K = self._prep(xt_prob, x0hat_prob, t=t)
        out = torch.empty_like(xt)
        L.check(L.lib.fddm_q_posterior_dense(xt.data_ptr(), xh.data_ptr(), t.data_ptr(), self.betas.data_ptr(),
                                             self.T, B, Lq, K, self.eps, out.data_ptr(), L.stream_ptr(xt.device)),
                "q_posterior")
        return out

    @torch.no_grad()
    def multistep_coeffs(self, t: torch.Tensor, delta: int, *, delta_cap: Optional[torch.Tensor] = None):
        """Per-sample coefficients of the Delta-step transition product (sched:132-183) computed on
        the device without host synchronisation: fp32 [4B+1] = a_cum | b_cum | a_tgt | b_tgt | identity.
        `delta_cap`: optional int64 device scalar, the global-batch min of t when the batch is
        sharded over ranks (quirk Q10: the reference clips delta by t.min() over the whole batch)."""
        return multistep_coeffs(t.to(self.device), int(delta), self.betas, self.K, delta_cap=delta_cap)

    @torch.no_grad()
    def q_posterior_multi_step(self, xt_prob: torch.Tensor, x0hat_prob: torch.Tensor, t: torch.Tensor,
                               delta: int) -> torch.Tensor:
        """Delta-step posterior q(x_{t-Delta}|x_t, x0hat).  sched:106-208.  The reference clips
        `delta` with `t.min().item()` (a host sync) and loops over the batch in Python; here both
        happen in one tiny kernel.  When the clipped delta is <= 0 the result equals `xt_prob`
        (the reference returns the input tensor itself; this returns a copy)."""
        (xt, xh), t, B, Lq, K = self._prep(xt_prob, x0hat_prob, t=t)
        coeffs = self.multistep_coeffs(t, int(delta))
        out = torch.empty_like(xt)
        L.check(L.lib.fddm_q_posterior_multi_dense(xt.data_ptr(), xh.data_ptr(), coeffs.data_ptr(), B, Lq, K,
                                                   self.eps, out.data_ptr(), L.stream_ptr(xt.device)),
                "q_posterior_multi_step")
        return out

    @property
    def w_prefix(self):
        """w_t = prod_{s<=t}(1-beta_s) = alpha_bar.  sched:210-213."""
        return self.alpha_bar

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample_q_ids(self, x0: torch.Tensor, t: torch.Tensor, *, exp_noise: Optional[torch.Tensor] = None,
                     generator: Optional[torch.Generator] = None,
                     philox_state: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Fused ids -> ids forward corruption == one-hot -> q_sample -> torch.multinomial(.,1)
        (train.py:180-188) without materialising the [B,L,K] one-hot / probability tensors.
        `exp_noise`: optional injected Exp(1) variates fp32 [B,L,K] (parity tests); otherwise
        in-kernel Philox keyed by the torch CUDA generator's (seed, offset), or by `philox_state`, an
        int64[2] device tensor {seed, offset} read by the kernel (the caller advances the offset; this
        is what makes the call replayable inside a CUDA graph)."""
        if x0.dim() != 2:
            raise ValueError(f"x0 must be (B, L) token ids, got shape {tuple(x0.shape)}")
        x0 = x0.long().contiguous()
        t = t.to(x0.device).long().contiguous()
        dev = L.require_cuda(x0, t, self.alpha_bar)
        B, Lq = x0.shape
        if t.shape != (B,):
            raise ValueError(f"t must have shape ({B},), got {tuple(t.shape)}")
        seed = offset = 0
        if exp_noise is not None:
            if exp_noise.dtype != torch.float32 or exp_noise.numel() != B * Lq * self.K:
                raise ValueError("exp_noise must be float32 with B*L*K elements")
            exp_noise = exp_noise.contiguous()
            L.require_cuda(exp_noise, x0)
        elif philox_state is None:
            seed, offset = philox_seed_offset(dev, generator, 4)
        out = torch.empty_like(x0)
        L.check(L.lib.fddm_sample_q_ids(x0.data_ptr(), t.data_ptr(), self.alpha_bar.data_ptr(), self.T, B, Lq,
                                        self.K, self.eps, L.ptr(exp_noise), seed, offset, L.ptr(philox_state),
                                        out.data_ptr(), L.stream_ptr(dev)), "sample_q_ids")
        return out


@torch.no_grad()
def multistep_coeffs(t: torch.Tensor, delta: int, betas: torch.Tensor, K: int, *,
                     delta_cap: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fddm_multistep_coeffs for any scheduler object that exposes a `betas` table (sched:132-183)."""
    t = t.long().contiguous()
    betas = betas.to(device=t.device, dtype=torch.float32).contiguous()
    L.require_cuda(t, betas)
    B = t.numel()
    coeffs = torch.empty(4 * B + 1, dtype=torch.float32, device=t.device)
    L.check(L.lib.fddm_multistep_coeffs(t.data_ptr(), B, int(delta), L.ptr(delta_cap), betas.data_ptr(),
                                        int(betas.numel()), int(K), coeffs.data_ptr(), L.stream_ptr(t.device)),
            "multistep_coeffs")
    return coeffs


def philox_seed_offset(device: torch.device, generator: Optional[torch.Generator], increment: int):
    """(seed, offset) of the torch CUDA generator, advancing its offset so successive calls draw
    independent streams (same contract as ATen's philox_cuda_state)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = generator if generator is not None else torch.cuda.default_generators[idx]
    seed = int(gen.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    offset = int(gen.get_offset())
    gen.set_offset(offset + ((int(increment) + 3) // 4) * 4)
    return seed, offset & 0xFFFFFFFFFFFFFFFF
