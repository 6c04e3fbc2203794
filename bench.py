#!/usr/bin/env python
"""bench.py -- the token-path benchmark (BASELINE.json metric: posterior+KL+resample tokens*V/s and
% of HBM peak at 1/2/4/8 B200 next to the host CPU).

A "step" is one pass of the hot path over one batch of synthetic input, exactly the calls train.py /
the sampler make per batch:
    xt   = SchedulerAdapter.sample_q(x0, t)                       fused ids->ids, in-kernel Philox
    kl   = SchedulerAdapter.kl_term(xt, x0, logits, t, mask)      fused forward + d/dlogits, one pass
    x'   = DiffusionJumpySampler._jump_once(...)                  exact Delta-step posterior + Categorical
    lfd  = lfd_loss(z_speech, z_text, lambda)                     stats + tcgen05 cross-covariance
    (kl + tau * w_t * lfd).backward()                             L_fd backward contractions
Work unit = one token*vocab element; a step processes B*L*V of them per GPU.

Workload (config.workload): the per-GPU shard of BASELINE configs[4] (global B=512 L=256 V=8000 D=768
over 8 GPUs => B=64 per GPU); the global batch is 64*N, so N=8 is exactly configs[4] ("weak" scaling).
Its 524 MB of fp32 logits per step exceed the 126 MB L2, so no L2 flush is needed between iterations.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (N>1: launched by torchrun)
    python bench.py --impl reference ...                           CPU arm: the torch-eager port of the
                                                                   reference path on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "fddm-asr_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "posterior+KL+resample token*vocab elements/s"
UNIT = "G tok*V/s"
WORKLOADS = {
    # name: (B per GPU, L, V, D)
    "c5shard": (64, 256, 8000, 768),
    "c2": (32, 128, 8000, 768),
    "c4": (64, 256, 32000, 768),
    "c1": (8, 64, 4000, 256),          # BASELINE configs[0]: the reference's own CPU-sized case (used by the CPU tests)
}
T_TRAIN, T_INFER, R_JUMP, LAMBDA, TAU = 200, 20, 5, 5e-3, 1.0
CPU_SLICE_B = 8            # batch slice the CPU arm runs per step (the path is linear in B)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synth(B, L, V, D, seed, dtype, pin):
    """Seeded synthetic batch in host memory: N(0,1) logits, ids with a pad tail
    (~70 % coverage), t in 1..T, correlated features."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, L, V, generator=g, dtype=torch.float32).to(dtype)
    x0 = torch.randint(1, V, (B, L), generator=g)
    lens = torch.randint(int(0.4 * L), L + 1, (B,), generator=g)
    mask = torch.arange(L)[None, :] < lens[:, None]
    x0 = torch.where(mask, x0, torch.zeros_like(x0))              # pad_id = 0
    t = torch.randint(1, T_TRAIN + 1, (B,), generator=g)
    za = torch.randn(B, L, D, generator=g)
    zb = 0.9 * za + (1 - 0.81) ** 0.5 * torch.randn(B, L, D, generator=g)
    out = dict(logits=logits, x0=x0, mask=mask, t=t, za=za.to(dtype), zb=zb.to(dtype))
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the torch-eager port of the reference path (oracle/fddm_torch_port.py)
# ------------------------------------------------------------------------------------------------
def cpu_step(P, d, betas, abar, V):
    xt = P.sample_q(d["x0"], d["t"], abar, V)
    lg = d["logits"].float().requires_grad_(True)
    kl = P.kl_term(xt, d["x0"], lg, d["t"], betas, d["mask"])
    x_new, _ = P.jump_once(d["x0"], d["logits"].float(), T_INFER, R_JUMP, betas, abar, V, T_TRAIN, T_INFER,
                           sampling_mode="exact", greedy=False)
    za = d["za"].float().requires_grad_(True); zb = d["zb"].float().requires_grad_(True)
    lfd = P.lfd_loss(za, zb, LAMBDA)
    (kl + TAU * abar[d["t"] - 1].mean() * lfd).backward()
    return float(kl.detach()), x_new


def run_cpu(workload, steps, warmup, dtype):
    from oracle import fddm_torch_port as P
    B, L, V, D = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bs = min(CPU_SLICE_B, B)
    d = synth(Bs, L, V, D, 1337, torch.float32, pin=False)
    betas, abar = P.make_tables(V, T_TRAIN)
    for _ in range(warmup):
        cpu_step(P, d, betas, abar, V)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(P, d, betas, abar, V)
    dt = (time.perf_counter() - t0) / max(1, steps)
    val = Bs * L * V / dt / 1e9
    sample = (f"torch-eager port of the reference path (oracle/fddm_torch_port.py), batch slice B={Bs} of {B} "
              f"(L={L} V={V} D={D}), {steps} steps after {warmup} warm-up, fp32, per-element throughput")
    return val, dt * 1e3, cores, sample


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class ResidentDecoder:
    """The denoising decoder is out of scope: it is replaced by pre-generated logits."""

    def __init__(self):
        self.logits = None

    def __call__(self, x, t, c):
        return self.logits


def shard_check(fb, dev, group, world, rank, bn_allreduce="moment"):
    """world>1 only: the batch-sharded kl_term / lfd_loss (values AND gradients, through the same host classes
    the timed step uses) must equal this library's single-process evaluation of the whole batch (which the
    -m gpu parity tests pin to the oracle).  Small batch; every rank builds the same global batch."""
    dist = torch.distributed
    Bg, L, V, D, T = 8 * world, 32, 4000, 256, 200
    g = torch.Generator(device=dev).manual_seed(1)
    logits = torch.randn(Bg, L, V, generator=g, device=dev) * 2
    x0 = torch.randint(0, V, (Bg, L), generator=g, device=dev)
    xt = torch.where(torch.rand(Bg, L, generator=g, device=dev) < 0.5, x0, torch.randint(0, V, (Bg, L), generator=g, device=dev))
    t = torch.randint(1, T + 1, (Bg,), generator=g, device=dev)
    mask = torch.rand(Bg, L, generator=g, device=dev) < 0.7
    za = torch.randn(Bg, L, D, generator=g, device=dev)
    zb = 0.8 * za + 0.6 * torch.randn(Bg, L, D, generator=g, device=dev)
    sch = fb.DiscreteDiffusionScheduler(K=V, T=T, device=dev)
    sl = slice(rank * Bg // world, (rank + 1) * Bg // world)
    lg = logits.clone().requires_grad_(True); a = za.clone().requires_grad_(True); b = zb.clone().requires_grad_(True)
    kl_ref = fb.SchedulerAdapter(sch).kl_term(xt, x0, lg, t, mask)
    lf_ref = fb.lfd_loss(a, b, LAMBDA)
    (kl_ref + 0.5 * lf_ref).backward()
    lgs = logits[sl].clone().requires_grad_(True); a_s = za[sl].clone().requires_grad_(True); b_s = zb[sl].clone().requires_grad_(True)
    kl = fb.SchedulerAdapter(sch, group=group).kl_term(xt[sl], x0[sl], lgs, t[sl], mask[sl])
    op = fb.LfdPipeline(a_s, b_s, LAMBDA, group=group, bn_allreduce=bn_allreduce)
    op.stats(); op.xcov()
    lf = op.loss()
    (kl + 0.5 * lf).backward()
    part = fb.SchedulerAdapter(sch, group=group, defer_reduce=True).kl_term(xt[sl], x0[sl], logits[sl], t[sl], mask[sl]).detach().clone()
    dist.all_reduce(part, group=group)
    torch.cuda.synchronize(dev)

    def rel(x, y):
        return float((x.double() - y.double()).abs().max() / y.double().abs().max().clamp_min(1e-30))
    errs = {"kl": abs(float(kl.detach()) - float(kl_ref.detach())) / abs(float(kl_ref.detach())), "lfd": abs(float(lf.detach()) - float(lf_ref.detach())) / abs(float(lf_ref.detach())),
            "dlogits": rel(lgs.grad, lg.grad[sl]), "dza": rel(a_s.grad, a.grad[sl]), "dzb": rel(b_s.grad, b.grad[sl]),
            "kl_deferred": abs(float(part) - float(kl_ref.detach())) / abs(float(kl_ref.detach()))}
    worst = torch.tensor([max(errs.values())], device=dev, dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    tol = 2e-5
    out = {"ok": bool(float(worst) < tol), "tol": tol, "worst_over_ranks": float(worst), "bn_allreduce": bn_allreduce,
           "global_batch": Bg, **{k: float(f"{v:.3e}") for k, v in errs.items()}}
    return out


def run_gpu(args):
    import fddm_b200 as fb
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    if fb._lib.MISSING:
        raise SystemExit(f"libfddm_b200.so does not export {fb._lib.MISSING}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torchrun", file=sys.stderr)

    shard = None
    if world > 1:
        shard = shard_check(fb, dev, group, world, rank)
        if not shard["ok"]:
            if rank == 0:
                print(json.dumps({"error": "shard_check failed: the batch-sharded path disagrees with the whole-batch "
                                           "evaluation", "shard_check": shard, "n_gpus": world}), flush=True)
            teardown(world, dev, code=3)

    B, L, V, D = WORKLOADS[args.workload]
    dtype = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[args.dtype]
    s_bytes = 4 if args.dtype == "f32" else 2
    host = synth(B, L, V, D, 1337 + rank, dtype, pin=True)
    d = {k: v.to(dev) for k, v in host.items()}
    d["logits"].requires_grad_(True); d["za"].requires_grad_(True); d["zb"].requires_grad_(True)

    sch = fb.DiscreteDiffusionScheduler(K=V, T=T_TRAIN, device=dev)
    # multi-GPU: the KL scalar is folded into the one scalar all-reduce at the end of the step
    ad = fb.SchedulerAdapter(sch, group=group, defer_reduce=True)
    dec = ResidentDecoder()
    smp = fb.DiffusionJumpySampler(sch, dec, K=V, T_train=T_TRAIN, T_infer=T_INFER, r=R_JUMP, greedy=False,
                                   sampling_mode="exact", device=dev)
    cond = torch.zeros(B, 1, 1, device=dev)
    kl_ev = []
    # device-side Philox {seed, offset}: read by the kernels, advanced by a (captured) add each step, so
    # that a CUDA-graph replay draws fresh noise exactly like an eager step would
    pstate = torch.tensor([1337 + rank, 0], dtype=torch.int64, device=dev)
    smp.philox_state = pstate

    ext_ev = []                                                   # external events recorded inside the captured step

    def step(dd, timed_kl=False, capture=False):
        for k in ("logits", "za", "zb"):
            dd[k].grad = None
        # (overlapping the all-reduces with the persistent row kernels on a side stream was measured at N=8
        #  and is slower -- the collective's CTAs wait on peers while holding SMs -- so they stay in order)
        lfd_op = fb.LfdPipeline(dd["za"], dd["zb"], LAMBDA, group=group, overlap=False)
        lfd_op.stats()                                            # + all-reduce of the batch statistics
        pstate[1:].add_(8)
        xt = ad.sample_q(dd["x0"], dd["t"], philox_state=pstate)
        if timed_kl or capture:
            e0 = torch.cuda.Event(enable_timing=True, external=capture)
            e1 = torch.cuda.Event(enable_timing=True, external=capture)
            e0.record()
        kl = ad.kl_term(xt, dd["x0"], dd["logits"], dd["t"], dd["mask"])
        if timed_kl or capture:
            e1.record(); (ext_ev if capture else kl_ev).append((e0, e1))
        lfd_op.xcov(piggyback=kl if world > 1 else None)          # + all-reduce of the covariance (carrying the KL partial sum)
        dec.logits = dd["logits"].detach()
        x_new, _ = smp._jump_once(dd["x0"], T_INFER, R_JUMP, cond, L, want_p=False)
        lfd = lfd_op.loss()
        total = kl + TAU * ad.w_t(dd["t"]).mean() * lfd.float()
        total.backward()
        if world > 1:                                             # reported loss: global KL rode along with the covariance
            total = lfd_op.piggyback + (total.detach() - kl.detach())
        return total, x_new

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        return float(tt.item())

    # ---- device-resident throughput (`value`) ----------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                                            # samples cover warm-up + both timed regions
    for _ in range(args.warmup):
        step(d)
    sync_all()
    # The step is ~30 short launches (0.9 ms of device time, 0.6 ms of host time to enqueue -- more when 8
    # ranks share one host), so the whole step (forward, collectives, backward) is captured once in a CUDA
    # graph and replayed: every library call is stream-ordered and sync-free by contract.
    graph = None
    launches_per_step = None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):
                    step(d)
            torch.cuda.current_stream(dev).wait_stream(side)
            sync_all()
            graph = torch.cuda.CUDAGraph()
            n0 = fb._lib.launch_count()
            with torch.cuda.graph(graph):
                g_total, g_xnew = step(d, capture=True)           # external events bracket the KL kernel in the graph
            launches_per_step = fb._lib.launch_count() - n0
            graph.replay()
            sync_all()
        except Exception as e:                                    # report and fall back to eager launches
            if rank == 0:
                print(f"bench.py: CUDA-graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize(dev)
    n0 = fb._lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    h0 = time.perf_counter()
    kl_in_graph = []
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            total, x_new = step(d, timed_kl=True)
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps       # CPU time to enqueue one step (no sync)
    ev1.record()
    sync_all()
    if graph is not None:
        total, x_new = g_total, g_xnew
        launches = launches_per_step * args.steps
        try:                                                      # the KL kernel inside the last replay of the timed region
            kl_in_graph.append(ext_ev[0][0].elapsed_time(ext_ev[0][1]))
        except Exception:
            pass
        # the dominant kernel timed on its own stream position, eagerly, over the same number of launches
        # (back to back between one pair of events, so the queue never drains while the host prepares a call)
        xt_k = ad.sample_q(d["x0"], d["t"], philox_state=pstate)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ad.kl_term(xt_k, d["x0"], d["logits"], d["t"], d["mask"])
        e1.record()
        torch.cuda.synchronize(dev)
        kl_batch_ms = e0.elapsed_time(e1) / args.steps
        roof_extra = {}
    else:
        launches = fb._lib.launch_count() - n0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    if graph is None:
        kl_ms, kl_how = sum(a.elapsed_time(b) for a, b in kl_ev) / max(1, len(kl_ev)), "CUDA events around the launch in every timed step"
    elif kl_in_graph and kl_in_graph[0] > 0:
        kl_ms, kl_how = kl_in_graph[0], "external CUDA events captured around the launch, last replay of the timed region"
    else:
        kl_ms, kl_how = kl_batch_ms, "CUDA events around K back-to-back eager launches"
    elems = B * L * V                                            # per GPU per step
    ms_step = ms_total / args.steps
    value = world * elems / (ms_step * 1e-3) / 1e9

    # ---- end to end through the host API with HOST buffers (`e2e`) ------------------------------
    dh = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    for k in ("logits", "za", "zb"):
        dh[k].requires_grad_(True)
    res_host = torch.empty(1, dtype=torch.float32).pin_memory()
    ids_host = torch.empty(B, L, dtype=torch.int64).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = res_host.numel() * 4 + ids_host.numel() * 8

    def e2e_step():
        with torch.no_grad():
            for k, v in host.items():
                dh[k].copy_(v, non_blocking=True)
        total, x_new = step(dh)
        res_host.copy_(total.detach().reshape(1), non_blocking=True)
        ids_host.copy_(x_new, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()              # the training loop's .item()

    for _ in range(min(2, args.warmup)):
        e2e_step()
    sync_all()
    e2e_steps = max(3, min(args.steps, 10))
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(e2e_steps):
        e2e_step()
    ee1.record()
    sync_all()
    clk = clocks.stop() if rank == 0 else {}
    e2e_ms = max_over_ranks(ee0.elapsed_time(ee1)) / e2e_steps
    e2e_val = world * elems / (e2e_ms * 1e-3) / 1e9

    if graph is None:
        roof_extra = {}
    # ---- roofline of the dominant kernel: fused KL forward+backward (2*s bytes per element) ------
    peak, peak_src = peaks()
    # logits of masked rows are never read (their gradient rows are still written): count what the
    # algorithm must move -- read s bytes per element of the valid rows, write s bytes per element of all rows
    valid_rows = int(host["mask"].sum())
    algo_bytes = float(s_bytes) * V * (valid_rows + B * L)
    achieved = algo_bytes / (kl_ms * 1e-3) / 1e9
    traffic = None                                                # dram bytes per launch from the committed ncu capture
    try:
        if args.workload == "c5shard" and args.dtype == "f32":
            with open(os.path.join(ROOT, "profiles", "r01b_kl_traffic.json")) as f:
                traffic = float(json.load(f)["traffic_bytes_per_launch"])
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "kernel": "kl_rows_ring_kernel (fused KL forward+backward)", "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "traffic_source": "profiles/r01b_kl_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch)" if traffic else None,
                "algorithmic_bytes_per_launch": algo_bytes, "valid_row_fraction": round(valid_rows / (B * L), 4), "ms_per_launch": round(kl_ms, 4), "timed_by": kl_how, "peak_source": peak_src,
                "frac_of_nominal_8TBs": round(achieved / 8000.0, 4), **roof_extra}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, ms, cores, sample = run_cpu(args.workload, 3, 1, "f32")
        cpu = {"value": round(v, 5), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "ms_per_step_slice": round(ms, 1), "cpu_model": cpu_model()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic (seeded N(0,1) logits, padded ids, correlated features)",
            "config": {"workload": f"{args.workload}: per-GPU shard B={B} L={L} V={V} D={D} of BASELINE configs[4] "
                                   f"(global B={B * world}); step = sample_q + KL fwd+bwd + exact jump/resample + L_fd fwd+bwd",
                       "global_batch": B * world, "seq_len": L, "vocab": V, "d_proj": D, "T": T_TRAIN,
                       "parallelism": f"batch-sharded x{world}" + (" (NCCL all-reduce: KL scalar, L_fd stats/cov/bn sums)" if world > 1 else ""),
                       "l2": f"inputs {h2d / 1e6:.0f} MB/step > 126 MB L2 (no flush needed)"},
            "e2e": {"value": round(e2e_val, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_ms, 4),
            "cuda_graph": graph is not None,
            "clocks": clk,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "loss": float(total.detach()),
            "shard_check": shard,
        }
        print(json.dumps(line), flush=True)
    # the captured graph holds NCCL kernels (a reference on the communicator): it must die before the group does
    graph = g_total = g_xnew = total = x_new = None
    ext_ev.clear(); kl_ev.clear()
    teardown(world, dev)


def teardown(world, dev, code=0):
    """Leave the process without hanging.  Round-1 incident: every multi-rank run hung at
    exit once the step was a CUDA graph, because destroy_process_group() was called while the captured graph --
    which holds NCCL kernels and so a reference on the communicator -- was still alive.  Order here: the JSON
    line is already printed and flushed; drop the graph and everything captured with it, synchronise, barrier,
    then destroy the group under a watchdog, and finally _exit so that no destructor can block either."""
    import gc
    import threading
    sys.stdout.flush(); sys.stderr.flush()
    if world <= 1:
        if code:
            sys.exit(code)
        return
    wd = threading.Timer(45.0, lambda: os._exit(code))            # the result is out: never hold the GPUs hostage
    wd.daemon = True
    wd.start()
    gc.collect()
    torch.cuda.synchronize(dev)
    try:
        torch.distributed.barrier()
        torch.cuda.synchronize(dev)
        torch.distributed.destroy_process_group()
    except Exception as e:                                      # pragma: no cover
        print(f"bench.py: teardown: {type(e).__name__}: {e}", file=sys.stderr)
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(code)


def run_reference(args):
    """`--impl reference`: the reference path's CPU implementation (torch-eager port; the reference is
    Python and cannot be installed on the GPU box) on the host cores, same metric/unit/config.
    Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    B, L, V, D = WORKLOADS[args.workload]
    v, ms, cores, sample = run_cpu(args.workload, args.steps, args.warmup, "f32")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 5), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (same generator as the B200 arm)",
        "config": {"workload": f"{args.workload}: per-GPU shard B={B} L={L} V={V} D={D} of BASELINE configs[4]; each step is a "
                               f"bounded sample (batch slice B={min(CPU_SLICE_B, B)}) of it",
                   "seq_len": L, "vocab": V, "d_proj": D, "T": T_TRAIN, "parallelism": f"host CPU, {cores} threads"},
        "cpu_baseline": {"value": round(v, 5), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model()},
        "e2e": {"value": round(v, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5shard", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
