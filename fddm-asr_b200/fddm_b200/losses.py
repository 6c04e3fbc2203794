"""lfd_loss -- B200 host mirror of losses/fddm_losses.py:18-58 (cross-modal decorrelation loss).

    za~, zb~ = standardise(z_a), standardise(z_b)      along the batch axis per (position, channel)
    C = za~^T zb~ / (B*T)                              D x D, rows flattened over (b, t)
    loss = sum_j (1 - C_jj)^2 + lambda * sum_{j != k} C_jk^2

Same signature as the reference (`lfd_loss(z_a, z_b, lambda_offdiag=5e-3, eps=1e-5)`), differentiable
w.r.t. both inputs, result in the input dtype.  The D x D cross-covariance and the two backward
contractions run on tcgen05 tensor cores (fddm_lfd_xcov / fddm_lfd_backward); the batch statistics are
fp64 sums so a batch-sharded caller (`group=`) can all-reduce them: statistics, the partial
covariance and the batch-norm backward sums are the only three collectives.

The collectives are ncclAllReduce (`collective="nccl"`) or the library's own all-reduce kernel on buffers in
symmetric memory (`collective="p2p"` / `"nvls"`, `fddm_xgpu_allreduce`: 21-35 us instead of 36-63 us on 8 B200s).

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


class _SymmExchange:
    """Symmetric-memory exchange buffers of one process group: the three L_fd exchange
    arrays live in `torch.distributed._symmetric_memory` allocations (every rank has every peer's copy mapped over
    NVLink / NVSwitch) and are summed over the ranks IN PLACE by the library's own kernel (`fddm_xgpu_allreduce`:
    peer loads / stores of a 1/world slice, or multimem.ld_reduce / multimem.st through the switch) instead of
    ncclAllReduce.  Buffers are cached per (name, length, dtype) and reused by every step; `gen` counts the
    re-uses so that a backward whose saved statistics were overwritten by a later forward fails loudly."""

    _by_group = {}

    def __init__(self, group):
        import torch.distributed._symmetric_memory as symm
        self.symm = symm
        self.group = group
        self.group_name = group.group_name
        self.rank = torch.distributed.get_rank(group)
        self.world = torch.distributed.get_world_size(group)
        self.bufs = {}
        self.gen = {}
        self.max_ctas = 0           # 0: library default; set to the reserved SM count when overlapping (set_sm_reserve)

    @classmethod
    def of(cls, group) -> "_SymmExchange":
        ex = cls._by_group.get(group.group_name)
        if ex is None:
            ex = cls._by_group[group.group_name] = cls(group)
        return ex

    def buffer(self, name: str, numel: int, dtype: torch.dtype, dev: torch.device):
        """(tensor of `numel` elements, key); the allocation is padded to a multiple of 4 elements (zero pad).
        The first call for a key is COLLECTIVE (rendezvous) and must not happen inside a stream capture."""
        key = (name, int(numel), dtype, dev.index)
        ent = self.bufs.get(key)
        if ent is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("lfd_loss(collective='p2p'|'nvls'): run one eager step before capturing a CUDA graph "
                                   "(the symmetric buffers are allocated collectively on first use)")
            padded = (int(numel) + 3) // 4 * 4
            t = self.symm.empty(padded, dtype=dtype, device=dev)
            t.zero_()
            hdl = self.symm.rendezvous(t, self.group_name)
            if self.world not in (2, 4, 8) and not int(hdl.multicast_ptr):
                raise RuntimeError("the peer-to-peer all-reduce is built for 2, 4 or 8 ranks")
            if int(hdl.signal_pad_size) < int(L.lib.fddm_xgpu_signal_pad_bytes()):
                raise RuntimeError("symmetric-memory signal pad is too small")
            ent = self.bufs[key] = (t, hdl, padded)
            self.gen[key] = 0
        self.gen[key] += 1
        return ent[0][:numel], key

    def all_reduce(self, key, algo: str = "p2p"):
        t, hdl, padded = self.bufs[key]
        mc = int(hdl.multicast_ptr)
        code = L.XGPU_NVLS if (algo == "nvls" and mc) or self.world not in (2, 4, 8) else L.XGPU_P2P
        L.check(L.lib.fddm_xgpu_allreduce(int(hdl.buffer_ptrs_dev), mc or None, int(hdl.signal_pad_ptrs_dev),
                                          self.rank, self.world, t.element_size(), padded, code, int(self.max_ctas),
                                          L.stream_ptr(t.device)), "xgpu_allreduce")


def symmetric_exchange_available(group=None, multicast: bool = False) -> bool:
    """True when every rank of `group` can allocate and rendezvous symmetric memory -- with `multicast=True`, also
    with a multicast (NVLS) mapping -- i.e. when `lfd_loss(..., collective="p2p")` (resp. "nvls") can be used.
    Collective call: every rank must make it."""
    if not torch.distributed.is_initialized():
        return False
    group = group or torch.distributed.group.WORLD
    if torch.distributed.get_world_size(group) < 2:
        return False
    dev = torch.device("cuda", torch.cuda.current_device())

    def all_ranks(ok: int) -> bool:
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=group)
        return bool(int(flag))

    # two steps, so that a rank that cannot even allocate never leaves the others waiting in the rendezvous
    ok = 1
    try:
        import torch.distributed._symmetric_memory as symm
        symm.empty(4, dtype=torch.float32, device=dev)
    except Exception:                                            # noqa: BLE001 -- any failure means "not available"
        ok = 0
    if not all_ranks(ok):
        return False
    try:
        ex = _SymmExchange.of(group)
        _, key = ex.buffer("probe", 4, torch.float32, dev)
        if multicast and not int(ex.bufs[key][1].multicast_ptr):
            ok = 0
    except Exception:                                            # noqa: BLE001
        ok = 0
    return all_ranks(ok)


def _comm_stream(dev: torch.device) -> torch.cuda.Stream:
    s = _comm_streams.get(dev.index)
    if s is None:
        # high priority: the collective's few CTAs are placed ahead of the queued CTAs of the persistent
        # row kernels (whose dynamic row scheduler simply runs with fewer resident CTAs meanwhile)
        s = _comm_streams[dev.index] = torch.cuda.Stream(device=dev, priority=-1)
    return s


class _AsyncAllReduce:
    """SUM all-reduce of `x` on the side stream; `wait()` makes the current stream wait for it.
    With group None both calls are no-ops.  `symm` = (exchange, key): `x` is that symmetric buffer and the
    library's own kernel (`algo`: "p2p" or "nvls") does the reduction instead of NCCL."""

    def __init__(self, x: torch.Tensor, group, overlap: bool, symm=None, algo: str = "p2p"):
        self.done = None
        if group is None:
            return

        def reduce():
            if symm is not None:
                symm[0].all_reduce(symm[1], algo)
            else:
                torch.distributed.all_reduce(x, op=torch.distributed.ReduceOp.SUM, group=group)

        if not overlap:
            reduce()
            return
        dev = x.device
        cur = torch.cuda.current_stream(dev)
        comm = _comm_stream(dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        if symm is None and not torch.cuda.is_current_stream_capturing():   # (a captured graph keeps its pool alive)
            x.record_stream(comm)
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            reduce()
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
        ctx.meta = (B, T, D, op.dt, op.world, op.eps, op.group, op.private_ws, op.exchange, op.sums_gen, op.collective)
        return loss.to(z_a.dtype)                              # the reference's result has the input dtype

    @staticmethod
    def backward(ctx, grad_out):
        z_a, z_b, sums, G, ws = ctx.saved_tensors
        B, T, D, dt, world, eps, group, private_ws, exchange, sums_gen, collective = ctx.meta
        dev = z_a.device
        if exchange is not None and exchange.gen[sums_gen[0]] != sums_gen[1]:
            raise RuntimeError("lfd_loss(collective='p2p'|'nvls'): the shared exchange buffer holding this call's batch "
                               "statistics was overwritten by a later lfd_loss forward of the same shape; run "
                               "backward before the next forward, or use collective='nccl'")
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
            if exchange is not None:
                xb, key = exchange.buffer("bn", 2 * T * D, torch.float32, dev)
                if parts > 1:
                    torch.sum(bn.view(2 * T, parts, D), dim=1, out=xb.view(2 * T, D))
                else:
                    xb.copy_(bn)
                exchange.all_reduce(key, collective)
                bn, parts = xb, 1
            else:
                if parts > 1:
                    bn = bn.view(2 * T, parts, D).sum(dim=1).contiguous()
                    parts = 1
                torch.distributed.all_reduce(bn, op=torch.distributed.ReduceOp.SUM, group=group)
        L.check(L.lib.fddm_lfd_backward(*args, bn.data_ptr(), parts, 1, dz_a.data_ptr(), dz_b.data_ptr(), st),
                "lfd_backward[1]")
        return dz_a, dz_b, None


class LfdPipeline:
    def __init__(self, z_a: torch.Tensor, z_b: torch.Tensor, lambda_offdiag: float = 5.0e-3, eps: float = 1e-5, *,
                 group=None, overlap: bool = False, collective: str = "nccl"):
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
        if collective not in ("nccl", "p2p", "nvls"):
            raise ValueError("collective must be 'nccl', 'p2p' or 'nvls'")
        # "p2p" / "nvls": the exchange buffers are symmetric memory, summed by the library's own kernel
        self.collective = collective
        self.exchange = _SymmExchange.of(group) if (collective != "nccl" and self.world > 1) else None
        self.sums_gen = None
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
        symm = None
        if self.exchange is not None:
            self.sums, key = self.exchange.buffer("sums", 4 * T * D, torch.float64, self.dev)
            self.sums_gen = (key, self.exchange.gen[key])
            symm = (self.exchange, key)
        else:
            self.sums = torch.empty(4 * T * D, dtype=torch.float64, device=self.dev)
        za, zb = self.z_a.detach(), self.z_b.detach()
        L.check(L.lib.fddm_lfd_stats(za.data_ptr(), zb.data_ptr(), self.dt, B, T, D, self.sums.data_ptr(),
                                     L.stream_ptr(self.dev)), "lfd_stats")
        self._ar_sums = _AsyncAllReduce(self.sums, self.group, self.overlap, symm, self.collective)
        return self

    def xcov(self, piggyback: torch.Tensor = None):
        """`piggyback`: optional fp32 scalar that rides along in the covariance all-reduce (e.g. the
        rank-partial KL loss of a deferred-reduce adapter); its global sum is `self.piggyback` after
        `loss()`."""
        if self.sums is None:
            self.stats()
        B, T, D = self.shape
        self._ar_sums.wait()
        symm = None
        if self.exchange is not None:
            self.cov, key = self.exchange.buffer("cov", D * D + 1, torch.float32, self.dev)
            symm = (self.exchange, key)
        else:
            self.cov = torch.empty(D * D + 1, dtype=torch.float32, device=self.dev)
        if piggyback is not None:
            self.cov[D * D:].copy_(piggyback.detach().reshape(1))
        else:
            self.cov[D * D:].zero_()
        za, zb = self.z_a.detach(), self.z_b.detach()
        L.check(L.lib.fddm_lfd_xcov(za.data_ptr(), zb.data_ptr(), self.dt, B, T, D, self.sums.data_ptr(),
                                    float(B * self.world), self.eps, self.ws.data_ptr(), self.cov.data_ptr(),
                                    L.stream_ptr(self.dev)), "lfd_xcov")
        self._ar_cov = _AsyncAllReduce(self.cov, self.group, self.overlap, symm, self.collective)
        return self

    def loss(self) -> torch.Tensor:
        if self.cov is None:
            self.xcov()
        self._ar_cov.wait()
        self.piggyback = self.cov[-1]
        return _LfdLossFn.apply(self.z_a, self.z_b, self)


def lfd_loss(z_a: torch.Tensor, z_b: torch.Tensor, lambda_offdiag: float = 5.0e-3, eps: float = 1e-5, *,
             group=None, collective: str = "nccl") -> torch.Tensor:
    """`group`: optional torch.distributed process group over which the batch axis is sharded; the
    loss (and its gradients) are then those of the reference evaluated on the global batch.
    `collective`: "nccl" (ncclAllReduce), or "p2p" / "nvls": symmetric-memory buffers summed by the library's own
    `fddm_xgpu_allreduce` kernel (peer loads/stores, or in-switch multicast reduction); both need
    `symmetric_exchange_available(group)`."""
    return LfdPipeline(z_a, z_b, lambda_offdiag, eps, group=group, overlap=False, collective=collective).loss()
