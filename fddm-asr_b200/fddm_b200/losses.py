"""lfd_loss -- B200 host mirror of losses/fddm_losses.py:18-58 (cross-modal decorrelation loss).

    za~, zb~ = standardise(z_a), standardise(z_b)      along the batch axis per (position, channel)
    C = za~^T zb~ / (B*T)                              D x D, rows flattened over (b, t)
    loss = sum_j (1 - C_jj)^2 + lambda * sum_{j != k} C_jk^2

Same signature as the reference (`lfd_loss(z_a, z_b, lambda_offdiag=5e-3, eps=1e-5)`), differentiable
w.r.t. both inputs, result in the input dtype.  The D x D cross-covariance and the two backward
contractions run on tcgen05 tensor cores (fddm_lfd_xcov / fddm_lfd_backward); the batch statistics are
fp64 sums so a batch-sharded caller (`group=`) can all-reduce them: statistics, the partial
covariance and the batch-norm backward sums are the only three collectives.
"""
from __future__ import annotations

import torch

from . import _lib as L


def _all_reduce(x, group):
    if group is not None:
        torch.distributed.all_reduce(x, op=torch.distributed.ReduceOp.SUM, group=group)


class _LfdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_a, z_b, lambda_offdiag, eps, group):
        B, T, D = z_a.shape
        dev = z_a.device
        dt = L.dtype_code(z_a)
        world = 1 if group is None else torch.distributed.get_world_size(group)
        st = L.stream_ptr(dev)
        nbytes = int(L.lib.fddm_lfd_workspace_bytes(B, T, D))
        if any(ctx.needs_input_grad[:2]):
            # a workspace of its own, kept until backward: the packed tensor-core operand planes and the
            # standardisation tables written by the forward are reused by the backward contractions
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            ws[:256].zero_()
        else:
            ws = L.zeroed_workspace(dev, "lfd", nbytes)
        sums = torch.empty(4 * T * D, dtype=torch.float64, device=dev)
        L.check(L.lib.fddm_lfd_stats(z_a.data_ptr(), z_b.data_ptr(), dt, B, T, D, sums.data_ptr(), st), "lfd_stats")
        _all_reduce(sums, group)
        cov = torch.empty(D * D, dtype=torch.float32, device=dev)
        L.check(L.lib.fddm_lfd_xcov(z_a.data_ptr(), z_b.data_ptr(), dt, B, T, D, sums.data_ptr(), float(B * world),
                                    float(eps), ws.data_ptr(), cov.data_ptr(), st), "lfd_xcov")
        _all_reduce(cov, group)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        G = torch.empty(D * D, dtype=torch.float32, device=dev)
        L.check(L.lib.fddm_lfd_loss(cov.data_ptr(), D, float(B * T * world), float(lambda_offdiag), ws.data_ptr(),
                                    loss.data_ptr(), G.data_ptr(), st), "lfd_loss")
        ctx.save_for_backward(z_a, z_b, sums, G, ws)
        ctx.meta = (B, T, D, dt, world, float(eps), group)
        return loss.to(z_a.dtype)                              # the reference's result has the input dtype

    @staticmethod
    def backward(ctx, grad_out):
        z_a, z_b, sums, G, ws = ctx.saved_tensors
        B, T, D, dt, world, eps, group = ctx.meta
        dev = z_a.device
        st = L.stream_ptr(dev)
        bn = torch.empty(4 * T * D, dtype=torch.float64, device=dev)
        dz_a = torch.empty_like(z_a)
        dz_b = torch.empty_like(z_b)
        g = grad_out.to(torch.float32).contiguous()
        args = (z_a.data_ptr(), z_b.data_ptr(), dt, B, T, D, sums.data_ptr(), float(B * world), eps, G.data_ptr(),
                float(B * T * world), g.data_ptr(), ws.data_ptr(), bn.data_ptr())
        L.check(L.lib.fddm_lfd_backward(*args, 0 | L.LFD_PLANES_VALID, dz_a.data_ptr(), dz_b.data_ptr(), st),
                "lfd_backward[0]")
        _all_reduce(bn, group)
        L.check(L.lib.fddm_lfd_backward(*args, 1, dz_a.data_ptr(), dz_b.data_ptr(), st), "lfd_backward[1]")
        return dz_a, dz_b, None, None, None


def lfd_loss(z_a: torch.Tensor, z_b: torch.Tensor, lambda_offdiag: float = 5.0e-3, eps: float = 1e-5, *,
             group=None) -> torch.Tensor:
    """`group`: optional torch.distributed process group over which the batch axis is sharded; the
    loss (and its gradients) are then those of the reference evaluated on the global batch."""
    if z_a.dim() != 3:
        raise ValueError(f"z_a must be (B, T, D), got shape {tuple(z_a.shape)}")
    B, T, D = z_a.shape
    assert z_b.shape == (B, T, D), "z_b must have the same shape as z_a"      # losses:36
    L.require_cuda(z_a, z_b)
    L.dtype_code(z_a)
    if z_b.dtype != z_a.dtype:
        z_b = z_b.to(z_a.dtype)
    return _LfdFn.apply(z_a.contiguous(), z_b.contiguous(), float(lambda_offdiag), float(eps), group)
