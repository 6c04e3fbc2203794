"""SchedulerAdapter -- the reference's designated plug-in seam (train.py:176-273), B200-native.

`sample_q(x0, t)`, `kl_term(xt, x0, logits_x0, t, x_mask=None)`, `w_t(t)` and the `.sch`
attribute keep the reference's signatures; `train_one_epoch` / `evaluate_*` accept an instance
as their `scheduler` argument unchanged.  The diffusion-KL is one fused CUDA kernel
(forward + gradient in a single pass over the logits) behind a `torch.autograd.Function`.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib as L
from .scheduler import DiscreteDiffusionScheduler


def _as_mask(x_mask: Optional[torch.Tensor], B: int, Lq: int) -> Optional[torch.Tensor]:
    """bool -> uint8 view (non-zero = valid); any other dtype -> fp32 WEIGHTS, because the reference multiplies
    the per-token KL by `x_mask.float()` (train.py:250): a non-boolean mask weights the tokens."""
    if x_mask is None:
        return None
    if x_mask.shape != (B, Lq):
        raise ValueError(f"x_mask must have shape ({B}, {Lq}), got {tuple(x_mask.shape)}")
    if x_mask.dtype == torch.bool:
        return x_mask.contiguous().view(torch.uint8)
    return x_mask.to(torch.float32).contiguous()


class _KLTermFn(torch.autograd.Function):
    """loss = mean_b [ sum_l m_bl KL_bl / (sum_l m_bl + eps) ]   (train.py:190-255).
    Differentiable w.r.t. logits only.  When a gradient is needed the forward launch already
    writes d loss / d logits (assuming an upstream gradient of `grad_scale`, default 1); backward
    rescales it in place only if the actual upstream gradient differs."""

    @staticmethod
    def forward(ctx, logits, xt, x0, t, mask, betas, T, batch_div, grad_scale, group, clamp_t):
        B, Lq, V = logits.shape
        dev = logits.device
        ws = L.zeroed_workspace(dev, "kl", int(L.lib.fddm_kl_workspace_bytes(B, Lq)))
        loss = torch.empty((), dtype=torch.float32, device=dev)
        need_grad = ctx.needs_input_grad[0]
        dt = L.dtype_code(logits)
        mask_f32 = (L.KL_MASK_F32 if (mask is not None and mask.dtype == torch.float32) else 0) | \
                   (L.KL_CLAMP_T if clamp_t else 0)
        if need_grad:
            grad = torch.empty_like(logits)
            L.check(L.lib.fddm_kl_forward_backward(logits.data_ptr(), dt, xt.data_ptr(), x0.data_ptr(), t.data_ptr(),
                                                   L.ptr(mask), mask_f32, betas.data_ptr(), T, B, Lq, V, float(batch_div),
                                                   L.ptr(grad_scale), ws.data_ptr(), loss.data_ptr(),
                                                   grad.data_ptr(), L.stream_ptr(dev)), "kl_forward_backward")
            ctx.save_for_backward(grad, grad_scale if grad_scale is not None else torch.empty(0, device=dev))
            ctx.has_scale = grad_scale is not None
        else:
            L.check(L.lib.fddm_kl_forward(logits.data_ptr(), dt, xt.data_ptr(), x0.data_ptr(), t.data_ptr(),
                                          L.ptr(mask), mask_f32, betas.data_ptr(), T, B, Lq, V, float(batch_div),
                                          ws.data_ptr(), loss.data_ptr(), L.stream_ptr(dev)), "kl_forward")
        if group is not None:
            torch.distributed.all_reduce(loss, op=torch.distributed.ReduceOp.SUM, group=group)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        grad, scale = ctx.saved_tensors
        if getattr(ctx, "consumed", False):
            raise RuntimeError("kl_term: the fused gradient buffer was already consumed by a previous backward "
                               "(retain_graph / double backward is not supported)")
        ctx.consumed = True
        g = grad_out.to(torch.float32).contiguous()
        L.check(L.lib.fddm_scale_inplace(grad.data_ptr(), L.dtype_code(grad), grad.numel(), g.data_ptr(),
                                         scale.data_ptr() if ctx.has_scale else None, L.stream_ptr(grad.device)),
                "scale_inplace")
        return grad, None, None, None, None, None, None, None, None, None, None


class SchedulerAdapter:
    def __init__(self, scheduler: DiscreteDiffusionScheduler, *, grad_scale=None, group=None,
                 defer_reduce: bool = False, validate_t: bool = True):
        """`grad_scale`: optional fp32 device scalar (or a zero-argument callable returning one),
        the upstream gradient the training loop will feed into `kl_term` (e.g. the AMP GradScaler's
        scale); folding it into the fused pass avoids a second pass over the gradient.
        `group`: optional torch.distributed process group; when given the batch is taken to be
        sharded over its ranks, the batch mean uses the global batch size and the scalar loss is
        all-reduced (SUM) -- the only collective of the KL path.  `defer_reduce=True` skips that
        all-reduce: kl_term then returns this rank's partial sum / global batch (its gradient is already
        the global loss's gradient), for callers that fold the scalar into a later collective."""
        self.sch = scheduler
        self._grad_scale = grad_scale
        self._group = group
        self._defer = bool(defer_reduce)
        self._validate_t = bool(validate_t)     # in-kernel trap on t outside 1..T (False: clamp silently)

    # -- train.py:180-188 ------------------------------------------------------------------------
    def sample_q(self, x0: torch.Tensor, t: torch.Tensor, *, exp_noise=None, generator=None,
                 philox_state=None) -> torch.Tensor:
        return self.sch.sample_q_ids(x0, t, exp_noise=exp_noise, generator=generator, philox_state=philox_state)

    # -- train.py:190-255 ------------------------------------------------------------------------
    def kl_term(self, xt: torch.Tensor, x0: torch.Tensor, logits_x0: torch.Tensor, t: torch.Tensor,
                x_mask: torch.Tensor = None) -> torch.Tensor:
        if not hasattr(self.sch, "betas"):
            raise ValueError("Scheduler must provide self.betas (shape [T]) to compute the posterior.")
        if logits_x0.dim() != 3:
            raise ValueError(f"logits_x0 must be (B, L, V), got shape {tuple(logits_x0.shape)}")
        B, Lq, V = logits_x0.shape
        dev = L.require_cuda(logits_x0)
        L.dtype_code(logits_x0)
        logits = logits_x0 if logits_x0.is_contiguous() else logits_x0.contiguous()
        xt = xt.to(dev).long().contiguous()
        x0 = x0.to(dev).long().contiguous()
        t = t.to(dev).long().contiguous()
        if xt.shape != (B, Lq) or x0.shape != (B, Lq) or t.shape != (B,):
            raise ValueError("xt/x0 must be (B, L) and t must be (B,)")
        betas = self.sch.betas.to(dev)
        if betas.dtype != torch.float32 or not betas.is_contiguous():
            betas = betas.float().contiguous()
        mask = _as_mask(None if x_mask is None else x_mask.to(dev), B, Lq)
        world = 1 if self._group is None else torch.distributed.get_world_size(self._group)
        gs = self._grad_scale() if callable(self._grad_scale) else self._grad_scale
        if gs is not None:
            gs = gs.detach().to(device=dev, dtype=torch.float32).reshape(())
        # validate_t: the kernel itself traps on a t outside 1..T (the reference raises: IndexError on CPU, device-side
        # assert on CUDA) -- no host synchronisation and no extra launch; validate_t=False clamps silently instead
        return _KLTermFn.apply(logits, xt, x0, t, mask, betas, int(betas.numel()), B * world, gs,
                               None if self._defer else self._group, not self._validate_t)

    # -- train.py:257-273 ------------------------------------------------------------------------
    def w_t(self, t: torch.Tensor) -> torch.Tensor:
        if hasattr(self.sch, "alpha_bar"):
            return self.sch.alpha_bar.to(t.device)[t - 1]
        if hasattr(self.sch, "w_prefix"):
            return self.sch.w_prefix.to(t.device)[t - 1]
        if hasattr(self.sch, "betas"):
            return torch.cumprod(1.0 - self.sch.betas.to(t.device), dim=0)[t - 1]
        return torch.ones_like(t, dtype=torch.float32)
