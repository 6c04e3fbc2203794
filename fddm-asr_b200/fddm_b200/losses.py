"""lfd_loss -- B200 host mirror of losses/fddm_losses.py:18-58 (cross-modal decorrelation loss).

    za~, zb~ = standardise(z_a), standardise(z_b)      along the batch axis per (position, channel)
    C = za~^T zb~ / (B*T)                              D x D, rows flattened over (b, t)
    loss = sum_j (1 - C_jj)^2 + lambda * sum_{j != k} C_jk^2

Same signature as the reference (`lfd_loss(z_a, z_b, lambda_offdiag=5e-3, eps=1e-5)`), differentiable
w.r.t. both inputs, result in the input dtype.  The D x D cross-covariance and the two backward
contractions run on tcgen05 tensor cores (fddm_lfd_xcov / fddm_lfd_backward); the batch statistics are
fp64 sums so a batch-sharded caller (`group=`) can all-reduce them: statistics, the partial
covariance and the batch-norm backward sums are the only three collectives.

`LfdPipeline` exposes the same computation in stages so that a batch-sharded training step can interleave
(or, with overlap=True, overlap on a high-priority side stream) the two forward all-reduces with the
independent KL / resampling kernels.  The persistent row kernels fill every SM, so a concurrent collective
must be given SMs of its own (`fddm_b200.set_sm_reserve(n)` with n = the collective's CTA count, e.g.
NCCL_MAX_CTAS); without that reservation the overlap LOSES (round 1, 8 B200s: the collective's CTAs queue
behind / spin next to the persistent CTAs):

    op = LfdPipeline(z_a, z_b, lam, eps, group=pg)
    op.stats()            # statistics kernel + async all-reduce
    ... launch KL ...
    op.xcov()             # waits for the statistics; pack + tcgen05 contraction + async all-reduce
    ... launch jump ...
    loss = op.loss()      # waits for the covariance; loss / dloss/dC; autograd node
"""
from __future__ import annotations

import torch

from . import _lib as L

_comm_streams = {}


def _comm_stream(dev: torch.device) -> torch.cuda.Stream:
    s = _comm_streams.get(dev.index)
    if s is None:
        # high priority: the collective's few CTAs are placed ahead of the queued CTAs of the persistent
        # row kernels (whose dynamic row scheduler simply runs with fewer resident CTAs meanwhile)
        s = _comm_streams[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
    return s


class _AsyncAllReduce:
    """SUM all-reduce of `x` on the side stream; `wait()` makes the current stream wait for it.
    With group None both calls are no-ops."""

    def __init__(self, x: torch.Tensor, group, overlap: bool):
        self.done = None
        if group is None:
            return
        if not overlap:
            torch.distributed.all_reduce(x, op=torch.distributed.ReduceOp.SUM, group=group)
            return
        dev = x.device
        cur = torch.cuda.current_stream(dev)
        comm = _comm_stream(dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        if not torch.cuda.is_current_stream_capturing():        # (a captured graph keeps its pool alive itself)
            x.record_stream(comm)
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            torch.distributed.all_reduce(x, op=torch.distributed.ReduceOp.SUM, group=group)
            self.done = torch.cuda.Event()
            self.done.record(comm)

    def wait(self):
        if self.done is not None:
            torch.cuda.current_stream().wait_event(self.done)
            self.done = None


class _LfdLossFn(torch.autograd.Function):
    """The last forward stage (loss + dloss/dC from the all-reduced covariance) and the backward."""

    @staticmethod
    def forward(ctx, z_a, z_b, op):
        B, T, D = z_a.shape
        dev = z_a.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        dv = op.d_valid
        if dv == D:
            G = torch.empty(D * D, dtype=torch.float32, device=dev)
            L.check(L.lib.fddm_lfd_loss(op.cov.data_ptr(), D, float(B * T * op.world), op.lam, op.ws.data_ptr(),
                                        loss.data_ptr(), G.data_ptr(), L.stream_ptr(dev)), "lfd_loss")
        else:
            # D was padded to a multiple of 8 for the tensor-core contraction: the loss and dloss/dC are taken over
            # the leading d_valid x d_valid block only; the padded columns have z~ = 0 and a zero row/column of G
            cov_v = op.cov[:D * D].view(D, D)[:dv, :dv].contiguous()
            G_v = torch.empty(dv * dv, dtype=torch.float32, device=dev)
            L.check(L.lib.fddm_lfd_loss(cov_v.data_ptr(), dv, float(B * T * op.world), op.lam, op.ws.data_ptr(),
                                        loss.data_ptr(), G_v.data_ptr(), L.stream_ptr(dev)), "lfd_loss")
            G = torch.zeros(D, D, dtype=torch.float32, device=dev)
            G[:dv, :dv] = G_v.view(dv, dv)
            G = G.view(-1)
        ctx.save_for_backward(z_a, z_b, op.sums, G, op.ws)
        ctx.meta = (B, T, D, op.dt, op.world, op.eps, op.group, op.private_ws)
        return loss.to(z_a.dtype)                              # the reference's result has the input dtype

    @staticmethod
    def backward(ctx, grad_out):
        z_a, z_b, sums, G, ws = ctx.saved_tensors
        B, T, D, dt, world, eps, group, private_ws = ctx.meta
        dev = z_a.device
        st = L.stream_ptr(dev)
        parts = int(L.lib.fddm_lfd_bn_parts(B, T, D))
        bn = torch.empty(2 * T * parts * D, dtype=torch.float32, device=dev)
        dz_a = torch.empty_like(z_a)
        dz_b = torch.empty_like(z_b)
        g = grad_out.to(torch.float32).contiguous()
        args = (z_a.data_ptr(), z_b.data_ptr(), dt, B, T, D, sums.data_ptr(), float(B * world), eps, G.data_ptr(),
                float(B * T * world), g.data_ptr(), ws.data_ptr())
        # the forward's packed operand planes / tables are reused when the workspace was private to this call
        phase0 = 0 | (L.LFD_PLANES_VALID if private_ws else 0)
        L.check(L.lib.fddm_lfd_backward(*args, bn.data_ptr(), parts, phase0, dz_a.data_ptr(), dz_b.data_ptr(), st),
                "lfd_backward[0]")
        if group is not None:
            # Only sum_b dz~*z~ crosses ranks, in fp32: [2][T][D].  (The other batch-norm moment, sum_b dz~ =
            # sum_k (sum_b z~[b,t,k]) G[.,k] / N, is identically zero over the GLOBAL batch because z~ has zero
            # global batch mean; the library never computes it.)
            if parts > 1:
                bn = bn.view(2 * T, parts, D).sum(dim=1).contiguous()
                parts = 1
            torch.distributed.all_reduce(bn, op=torch.distributed.ReduceOp.SUM, group=group)
        L.check(L.lib.fddm_lfd_backward(*args, bn.data_ptr(), parts, 1, dz_a.data_ptr(), dz_b.data_ptr(), st),
                "lfd_backward[1]")
        return dz_a, dz_b, None


class LfdPipeline:
    def __init__(self, z_a: torch.Tensor, z_b: torch.Tensor, lambda_offdiag: float = 5.0e-3, eps: float = 1e-5, *,
                 group=None, overlap: bool = False):
        if z_a.dim() != 3:
            raise ValueError(f"z_a must be (B, T, D), got shape {tuple(z_a.shape)}")
        B, T, D = z_a.shape
        assert z_b.shape == (B, T, D), "z_b must have the same shape as z_a"      # losses:36
        self.dev = L.require_cuda(z_a, z_b)
        self.dt = L.dtype_code(z_a)
        if z_b.dtype != z_a.dtype:
            z_b = z_b.to(z_a.dtype)
        # the reference accepts any D (losses:29-58); the contraction wants D % 8 == 0, so other widths are
        # zero-padded (a differentiable torch pad: the gradient is sliced back automatically)
        self.d_valid = D
        if D % 8:
            pad = 8 - D % 8
            z_a = torch.nn.functional.pad(z_a, (0, pad))
            z_b = torch.nn.functional.pad(z_b, (0, pad))
            D += pad
        self.z_a, self.z_b = z_a.contiguous(), z_b.contiguous()
        self.lam, self.eps, self.group = float(lambda_offdiag), float(eps), group
        self.world = 1 if group is None else torch.distributed.get_world_size(group)
        self.overlap = overlap
        self.shape = (B, T, D)
        nbytes = int(L.lib.fddm_lfd_workspace_bytes(B, T, D))
        self.private_ws = torch.is_grad_enabled() and (z_a.requires_grad or z_b.requires_grad)
        if self.d_valid != D and self.world > 1:
            raise ValueError("lfd_loss: D must be a multiple of 8 when the batch is sharded over ranks")
        if self.private_ws:
            # a workspace of its own, kept until backward: the packed tensor-core operand planes and the
            # standardisation tables written by the forward are reused by the backward contractions
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
            self.ws[:256].zero_()
        else:
            self.ws = L.zeroed_workspace(self.dev, "lfd", nbytes)
        self.sums = self.cov = None
        self._ar_sums = self._ar_cov = None

    def stats(self):
        B, T, D = self.shape
        self.sums = torch.empty(4 * T * D, dtype=torch.float64, device=self.dev)
        za, zb = self.z_a.detach(), self.z_b.detach()
        L.check(L.lib.fddm_lfd_stats(za.data_ptr(), zb.data_ptr(), self.dt, B, T, D, self.sums.data_ptr(),
                                     L.stream_ptr(self.dev)), "lfd_stats")
        self._ar_sums = _AsyncAllReduce(self.sums, self.group, self.overlap)
        return self

    def xcov(self, piggyback: torch.Tensor = None):
        """`piggyback`: optional fp32 scalar that rides along in the covariance all-reduce (e.g. the
        rank-partial KL loss of a deferred-reduce adapter); its global sum is `self.piggyback` after
        `loss()`."""
        if self.sums is None:
            self.stats()
        B, T, D = self.shape
        self._ar_sums.wait()
        self.cov = torch.empty(D * D + 1, dtype=torch.float32, device=self.dev)
        if piggyback is not None:
            self.cov[D * D:].copy_(piggyback.detach().reshape(1))
        else:
            self.cov[D * D:].zero_()
        za, zb = self.z_a.detach(), self.z_b.detach()
        L.check(L.lib.fddm_lfd_xcov(za.data_ptr(), zb.data_ptr(), self.dt, B, T, D, self.sums.data_ptr(),
                                    float(B * self.world), self.eps, self.ws.data_ptr(), self.cov.data_ptr(),
                                    L.stream_ptr(self.dev)), "lfd_xcov")
        self._ar_cov = _AsyncAllReduce(self.cov, self.group, self.overlap)
        return self

    def loss(self) -> torch.Tensor:
        if self.cov is None:
            self.xcov()
        self._ar_cov.wait()
        self.piggyback = self.cov[-1]
        return _LfdLossFn.apply(self.z_a, self.z_b, self)


def lfd_loss(z_a: torch.Tensor, z_b: torch.Tensor, lambda_offdiag: float = 5.0e-3, eps: float = 1e-5, *,
             group=None) -> torch.Tensor:
    """`group`: optional torch.distributed process group over which the batch axis is sharded; the
    loss (and its gradients) are then those of the reference evaluated on the global batch."""
    return LfdPipeline(z_a, z_b, lambda_offdiag, eps, group=group, overlap=False).loss()
