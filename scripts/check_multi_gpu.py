#!/usr/bin/env python
"""Multi-GPU correctness check (launch with torchrun, one rank per GPU, NCCL): the batch-sharded
kl_term / lfd_loss (values AND gradients) equal the single-process evaluation on the whole batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fddm-asr_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import fddm_b200 as fb


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Bg, L, V, D, T = 8 * world, 32, 4000, 256, 200
    g = torch.Generator(device="cuda").manual_seed(1)              # same global batch on every rank
    logits = torch.randn(Bg, L, V, generator=g, device=dev) * 2
    x0 = torch.randint(0, V, (Bg, L), generator=g, device=dev)
    xt = torch.where(torch.rand(Bg, L, generator=g, device=dev) < 0.5, x0, torch.randint(0, V, (Bg, L), generator=g, device=dev))
    t = torch.randint(1, T + 1, (Bg,), generator=g, device=dev)
    mask = torch.rand(Bg, L, generator=g, device=dev) < 0.7
    za = torch.randn(Bg, L, D, generator=g, device=dev)
    zb = 0.8 * za + 0.6 * torch.randn(Bg, L, D, generator=g, device=dev)
    sch = fb.DiscreteDiffusionScheduler(K=V, T=T, device=dev)
    sl = slice(rank * Bg // world, (rank + 1) * Bg // world)

    # whole batch, one process
    lg = logits.clone().requires_grad_(True); a = za.clone().requires_grad_(True); b = zb.clone().requires_grad_(True)
    kl_ref = fb.SchedulerAdapter(sch).kl_term(xt, x0, lg, t, mask)
    lf_ref = fb.lfd_loss(a, b, 5e-3)
    (kl_ref + 0.5 * lf_ref).backward()

    # sharded
    lgs = logits[sl].clone().requires_grad_(True); a_s = za[sl].clone().requires_grad_(True); b_s = zb[sl].clone().requires_grad_(True)
    kl = fb.SchedulerAdapter(sch, group=dist.group.WORLD).kl_term(xt[sl], x0[sl], lgs, t[sl], mask[sl])
    op = fb.LfdPipeline(a_s, b_s, 5e-3, group=dist.group.WORLD)
    op.stats(); op.xcov()
    lf = op.loss()
    (kl + 0.5 * lf).backward()
    torch.cuda.synchronize()

    def rel(x, y):
        return float((x.double() - y.double()).abs().max() / y.double().abs().max().clamp_min(1e-30))
    errs = {"kl": abs(float(kl) - float(kl_ref)) / abs(float(kl_ref)), "lfd": abs(float(lf) - float(lf_ref)) / abs(float(lf_ref)),
            "dlogits": rel(lgs.grad, lg.grad[sl]), "dza": rel(a_s.grad, a.grad[sl]), "dzb": rel(b_s.grad, b.grad[sl])}
    # deferred reduce: partial sums add up to the global loss
    part = fb.SchedulerAdapter(sch, group=dist.group.WORLD, defer_reduce=True).kl_term(xt[sl], x0[sl], logits[sl], t[sl], mask[sl]).detach().clone()
    dist.all_reduce(part)
    errs["kl_deferred"] = abs(float(part) - float(kl_ref)) / abs(float(kl_ref))
    ok = all(v < 2e-5 for v in errs.values())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi-gpu check", "OK" if int(flag) else "FAILED", "world", world, {k: f"{v:.2e}" for k, v in errs.items()})
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
