#!/usr/bin/env python
"""bench.py -- the token-path benchmark (BASELINE.json metric: posterior+KL+resample tokens*V/s and
% of HBM peak at 1/2/4/8 B200 next to the host CPU).

A "step" is one pass of the hot path over one batch of synthetic input, exactly the calls train.py /
the sampler make per batch.  Training-shaped workloads (c5, c5shard, c2, c4, c1):
    xt   = SchedulerAdapter.sample_q(x0, t)                       fused ids->ids, in-kernel Philox
    kl   = SchedulerAdapter.kl_term(xt, x0, logits, t, mask)      fused forward + d/dlogits, one pass
    x'   = DiffusionJumpySampler._jump_once(...)                  exact Delta-step posterior + Categorical
    lfd  = lfd_loss(z_speech, z_text, lambda)                     stats + tcgen05 cross-covariance  (D > 0)
    (kl + tau * w_t * lfd).backward()                             L_fd backward contractions
Sampler-shaped workload (c3): one `DiffusionJumpySampler.sample()` = ceil(T_infer / r) = 4 fused jumps.
Work unit = one token*vocab element; a training step processes B*L*V of them per GPU, a sampler step 4*B*L*V.

Workloads (config.workload) -- BASELINE.json configs[i]:
    c5       configs[4], THE DEFAULT: global B=512 L=256 V=8000 D=768, batch-sharded B=512/N per GPU, so N=1
             runs the full config on one GPU and N=8 is the 8xB200 case ("strong" scaling: fixed total work).
             This is what the driver's BENCH and SCALE runs measure.
    c5shard  the per-GPU shard of configs[4] at 8 ranks (B=64 per GPU, global 64*N; "weak" scaling)
    c2       configs[1]: B=32 L=128 V=8000 D=768 per GPU
    c3       configs[2]: B=256 L=128 V=8000, T_infer=20 r=5, --sampling-mode exact|fast, --greedy
    c4       configs[3]: B=64 L=256 V=32000 (KL + posterior/resample, no L_fd), --dtype f32|bf16
    c1       configs[0]: B=8 L=64 V=4000 (the reference's CPU-sized case; used by the CPU tests)
Inputs are far larger than the 126 MB L2 for every workload but c1, so no L2 flush is needed between steps.

    python bench.py [--gpus N] [--steps K] [--warmup W]            (N>1: launched by torchrun)
    python bench.py --impl reference ...                           CPU arm: the reference path on the host cores
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
# name: (B, L, V, D, kind, sharded)   sharded: B is the GLOBAL batch, split over the ranks
WORKLOADS = {
    "c5": (512, 256, 8000, 768, "train", True),
    "c5shard": (64, 256, 8000, 768, "train", False),
    "c2": (32, 128, 8000, 768, "train", False),
    "c3": (256, 128, 8000, 0, "sampler", False),
    "c4": (64, 256, 32000, 0, "train", False),
    "c1": (8, 64, 4000, 256, "train", False),
}
T_TRAIN, T_INFER, R_JUMP, LAMBDA, TAU = 200, 20, 5, 5e-3, 1.0
CPU_SLICE_B = 8            # batch slice the CPU arm runs per step (the path is linear in B)
EAGER_SLICE_B = 32         # batch slice of the torch-eager-on-B200 leg (its temporaries are ~15 x [B,L,V] fp32)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
                "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"}


def shape_of(workload, world):
    B, L, V, D, kind, sharded = WORKLOADS[workload]
    if sharded:
        if B % world:
            raise SystemExit(f"bench.py: workload {workload} needs the world size to divide B={B}")
        B //= world
    return B, L, V, D, kind, sharded


def synth(B, L, V, D, seed, dtype, device="cpu"):
    """Seeded synthetic batch: N(0,1) logits, ids with a pad tail (~70 % coverage),
    t in 1..T, correlated features."""
    g = torch.Generator(device=device).manual_seed(seed)
    logits = torch.randn(B, L, V, generator=g, dtype=torch.float32, device=device).to(dtype)
    x0 = torch.randint(1, V, (B, L), generator=g, device=device)
    lens = torch.randint(int(0.4 * L), L + 1, (B,), generator=g, device=device)
    mask = torch.arange(L, device=device)[None, :] < lens[:, None]
    x0 = torch.where(mask, x0, torch.zeros_like(x0))              # pad_id = 0
    t = torch.randint(1, T_TRAIN + 1, (B,), generator=g, device=device)
    out = dict(logits=logits, x0=x0, mask=mask, t=t)
    if D > 0:
        za = torch.randn(B, L, D, generator=g, device=device)
        zb = 0.9 * za + (1 - 0.81) ** 0.5 * torch.randn(B, L, D, generator=g, device=device)
        out.update(za=za.to(dtype), zb=zb.to(dtype))
    return out


# ------------------------------------------------------------------------------------------------
# baseline arms: the reference path in torch-eager form (the real reference when importable, else the
# port in oracle/fddm_torch_port.py), on the host cores or -- `eager_b200` -- on the GPU
# ------------------------------------------------------------------------------------------------
class _RefImpl:
    """The UNMODIFIED reference (fddm/sched/diffusion_scheduler.py, train.SchedulerAdapter,
    losses/fddm_losses.py, sampler/jumpy_sampler.py) behind the port's functional interface."""
    kind = "reference"

    def __init__(self, path):
        sys.dont_write_bytecode = True
        if path not in sys.path:
            sys.path.insert(0, path)
        import warnings
        warnings.filterwarnings("ignore")
        from fddm.sched.diffusion_scheduler import DiscreteDiffusionScheduler
        from losses.fddm_losses import lfd_loss
        from sampler.jumpy_sampler import DiffusionJumpySampler
        import train as ref_train
        self._S, self._lfd, self._J, self._A = DiscreteDiffusionScheduler, lfd_loss, DiffusionJumpySampler, ref_train.SchedulerAdapter
        self.path = path
        self._cache = {}

    def _sch(self, K, T, device):
        key = (K, T, str(device))
        if key not in self._cache:
            sch = self._S(K=K, T=T, device=device)
            self._cache[key] = (sch, self._A(sch))
        return self._cache[key]

    def make_tables(self, K, T, beta_max=0.2, device="cpu"):
        sch, _ = self._sch(K, T, torch.device(device))
        return sch.betas, sch.alpha_bar

    def sample_q(self, x0, t, alpha_bar, K):
        return self._sch(K, alpha_bar.numel(), x0.device)[1].sample_q(x0, t)

    def kl_term(self, xt, x0, logits, t, betas, x_mask=None):
        return self._sch(logits.shape[-1], betas.numel(), logits.device)[1].kl_term(xt, x0, logits, t, x_mask)

    def jump_once(self, x_t, logits, t_scalar, delta, betas, alpha_bar, K, T_train, T_infer, sampling_mode="exact",
                  greedy=True, temperature=1.0):
        sch, _ = self._sch(K, T_train, x_t.device)
        smp = self._J(sch, lambda x, tt, c: logits, K=K, T_train=T_train, T_infer=T_infer, r=delta, greedy=greedy,
                      sampling_mode=sampling_mode, temperature=temperature, device=x_t.device)
        return smp._jump_once(x_t, t_scalar, delta, torch.zeros(x_t.shape[0], 1, 1, device=x_t.device), x_t.shape[1])

    def lfd_loss(self, z_a, z_b, lambda_offdiag=5e-3, eps=1e-5):
        return self._lfd(z_a, z_b, lambda_offdiag, eps)


def load_baseline_impl():
    """The reference itself when it can be imported ($FDDM_REF, /root/reference, baseline/_ref), else the port
    (FDDM_BASELINE=port forces the port)."""
    force_port = os.environ.get("FDDM_BASELINE", "") == "port"
    for path in (os.environ.get("FDDM_REF"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if not force_port and path and os.path.exists(os.path.join(path, "fddm", "sched", "diffusion_scheduler.py")):
            try:
                return _RefImpl(path), f"the unmodified reference imported from {path}"
            except Exception as e:                                 # missing dependency etc.: say so, use the port
                print(f"bench.py: reference at {path} not importable ({type(e).__name__}: {e}); using the port",
                      file=sys.stderr)
    from oracle import fddm_torch_port as P
    P.kind = "port"
    return P, "torch-eager port of the reference path (oracle/fddm_torch_port.py)"


def baseline_step(P, d, betas, abar, V, kind, mode="exact", greedy=False):
    if kind == "sampler":
        B, L = d["x0"].shape
        x = torch.randint(0, V, (B, L), device=d["x0"].device)
        t = T_INFER
        while t > 0:
            delta = min(R_JUMP, t)
            x, p = P.jump_once(x, d["logits"].float(), t, delta, betas, abar, V, T_TRAIN, T_INFER, sampling_mode=mode,
                               greedy=greedy)
            t -= delta
        return 0.0, p.argmax(-1)
    xt = P.sample_q(d["x0"], d["t"], abar, V)
    lg = d["logits"].float().requires_grad_(True)
    kl = P.kl_term(xt, d["x0"], lg, d["t"], betas, d["mask"])
    x_new, _ = P.jump_once(d["x0"], d["logits"].float(), T_INFER, R_JUMP, betas, abar, V, T_TRAIN, T_INFER,
                           sampling_mode="exact", greedy=False)
    total = kl
    if "za" in d:
        za = d["za"].float().requires_grad_(True); zb = d["zb"].float().requires_grad_(True)
        total = kl + TAU * abar[d["t"] - 1].mean() * P.lfd_loss(za, zb, LAMBDA)
    total.backward()
    return float(kl.detach()), x_new


def run_baseline(workload, steps, warmup, device, slice_b, mode="exact", greedy=False):
    """Times `steps` baseline steps on a batch slice; returns (G tok*V/s, ms/step, description, kind)."""
    P, what = load_baseline_impl()
    B, L, V, D, kind, _ = WORKLOADS[workload]
    Bs = min(slice_b, B)
    d = synth(Bs, L, V, D, 1337, torch.float32, device=device)
    betas, abar = P.make_tables(V, T_TRAIN, device=device)
    sync = (lambda: torch.cuda.synchronize(device)) if str(device).startswith("cuda") else (lambda: None)
    for _ in range(warmup):
        baseline_step(P, d, betas, abar, V, kind, mode, greedy)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        baseline_step(P, d, betas, abar, V, kind, mode, greedy)
    sync()
    dt = (time.perf_counter() - t0) / max(1, steps)
    per_step = (4 if kind == "sampler" else 1) * Bs * L * V
    sample = (f"{what}, batch slice B={Bs} of {B} (L={L} V={V} D={D}), {steps} steps after {warmup} warm-up, fp32, "
              f"per-element throughput")
    return per_step / dt / 1e9, dt * 1e3, sample, P.kind


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
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Number of samples taken so far (to separate the load phases afterwards)."""
        try:
            with open(self.path) as f:
                return sum(1 for _ in f)
        except Exception:
            return 0

    def stop(self, lo=0, hi=None):
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
            lines = open(self.path).read().splitlines()
            for line in lines[lo:hi]:
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


def shard_check(fb, dev, group, world, rank, per_rank_b=8, collective="nccl"):
    """world>1 only: the batch-sharded kl_term / lfd_loss (values AND gradients, through the same host classes
    the timed step uses) must equal this library's single-process evaluation of the whole batch (which the
    -m gpu parity tests pin to the oracle).  Small batch; every rank builds the same global batch.
    per_rank_b = 8 exercises the small-batch L_fd path (B < 32 per rank), 32 the tb-major one."""
    dist = torch.distributed
    Bg, L, V, D, T = per_rank_b * world, 32, 4000, 256, 200
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
    op = fb.LfdPipeline(a_s, b_s, LAMBDA, group=group, collective=collective)
    op.stats(); op.xcov()
    lf = op.loss()
    (kl + 0.5 * lf).backward()
    part = fb.SchedulerAdapter(sch, group=group, defer_reduce=True).kl_term(xt[sl], x0[sl], logits[sl], t[sl], mask[sl]).detach().clone()
    dist.all_reduce(part, group=group)
    torch.cuda.synchronize(dev)

    def rel(x, y):
        return float((x.double() - y.double()).abs().max() / y.double().abs().max().clamp_min(1e-30))
    klr, lfr = float(kl_ref.detach()), float(lf_ref.detach())
    errs = {"kl": abs(float(kl.detach()) - klr) / abs(klr), "lfd": abs(float(lf.detach()) - lfr) / abs(lfr),
            "dlogits": rel(lgs.grad, lg.grad[sl]), "dza": rel(a_s.grad, a.grad[sl]), "dzb": rel(b_s.grad, b.grad[sl]),
            "kl_deferred": abs(float(part) - klr) / abs(klr)}
    worst = torch.tensor([max(errs.values())], device=dev, dtype=torch.float64)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX, group=group)
    tol = 2e-5
    return {"ok": bool(float(worst) < tol), "tol": tol, "worst_over_ranks": float(worst), "global_batch": Bg, **{k: float(f"{v:.3e}") for k, v in errs.items()}}


def kernel_roofline(prof, n_steps, shp, dtype_name, valid_rows, pk, sampler_jumps):
    """Per-kernel block from the library's own CUDA-event records: ms per launch, algorithmic bytes / useful
    flops per launch, achieved rate and fraction of the measured peak."""
    B, L, V, D = shp
    s = 4 if dtype_name == "f32" else 2
    rows, R = B * L, B * L                                         # token rows; feature rows (T == L here)
    planes = 2 if dtype_name == "bf16" else 4
    hbm = {
        "kl_rows_fwdbwd": s * V * (valid_rows + rows),             # read valid rows once + write every gradient row
        "kl_rows_fwd": s * V * valid_rows,
        "jump_rows_philox": s * V * rows, "jump_rows_greedy": s * V * rows,
        "jump_rows_injected_noise": (s + 4) * V * rows,
        "lfd_stats_kernel": 2 * R * D * s,
        "lfd_pack_kernel": 2 * R * D * s + planes * R * D * 2,
        "lfd_bn_reduce_kernel": 2 * R * D * (s + 4),
        "lfd_bn_finalize_kernel": 2 * R * D * (s + 4 + s),
    }
    if sampler_jumps:                                              # the last jump of a chain also writes p_x0
        for k in ("jump_rows_philox", "jump_rows_greedy"):
            hbm[k] = s * V * rows * (sampler_jumps + 1) / sampler_jumps
    out, lfd_ms, total_ms = [], 0.0, 0.0
    for name, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        per = ms / max(1, n)
        e = {"kernel": name, "launches_per_step": round(n / n_steps, 2), "ms": round(per, 4),
             "ms_per_step": round(ms / n_steps, 4)}
        total_ms += ms / n_steps
        if name.startswith("lfd_") or name.startswith("umma_"):
            lfd_ms += ms / n_steps
        if name.startswith("umma_"):
            fl = 2.0 * R * D * D
            tf = fl / (per * 1e-3) / 1e12
            terms = 1 if (planes == 2 and "fwd" in name) else 3
            e.update(bound="tensor", useful_flops=fl, achieved=round(tf, 1), unit="TFLOP/s", peak=pk["bf16_tflops"],
                     frac=round(tf / pk["bf16_tflops"], 4), issued_frac=round(terms * tf / pk["bf16_tflops"], 4),
                     note="issued = hi*hi + hi*lo + lo*hi bf16 MMAs per fp32-accurate product")
        elif hbm.get(name):
            by = float(hbm[name])
            gbs = by / (per * 1e-3) / 1e9
            e.update(bound="hbm", algorithmic_bytes=by, achieved=round(gbs, 1), unit="GB/s", peak=pk["hbm_gbs"],
                     frac=round(gbs / pk["hbm_gbs"], 4))
        out.append(e)
    return out, lfd_ms, total_ms


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
    # Running L_fd's forward exchanges under the KL / jump kernels costs those kernels the reserved SMs and hides a
    # roughly constant latency.  With NCCL (16 SMs = 11 % of the row kernels' time, ~90 us hidden) it pays below
    # roughly 40k token rows per GPU (N >= 4 on c5), not above (N = 2: 65k rows).  With the library's own exchange
    # kernel (8 SMs) it pays at every N measured (N = 2: 2.115 vs 2.137 ms/step).  Decided below, once the exchange
    # mode is known.
    rows_per_gpu = shape_of(args.workload, world)[0] * shape_of(args.workload, world)[1]
    nccl_ctas = int(os.environ.get("NCCL_MAX_CTAS", "16"))        # 4 / 8 / 16 / 24 / 32 measured at N=8: 16 is best
    if world > 1:
        import torch.distributed as dist
        if args.collectives != "serial":
            # an exchange that runs on a side stream under the row kernels is held to a few CTAs, and the persistent
            # row kernels leave exactly that many SMs free
            os.environ.setdefault("NCCL_MAX_CTAS", str(nccl_ctas))
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torchrun", file=sys.stderr)

    # How the three L_fd exchange buffers are summed over the ranks: ncclAllReduce, or the library's own kernel on
    # symmetric memory ("p2p": peer loads/stores; "nvls": in-switch multicast reduction).  "auto" takes p2p when
    # every rank can set up symmetric memory, NCCL otherwise.
    exchange = "nccl"
    if world > 1 and args.exchange != "nccl":
        if args.exchange == "auto":
            if fb.symmetric_exchange_available(group, multicast=True):
                exchange = "nvls"
            elif world in (2, 4, 8) and fb.symmetric_exchange_available(group):
                exchange = "p2p"
        elif not fb.symmetric_exchange_available(group, multicast=(args.exchange == "nvls")):
            raise SystemExit(f"bench.py: --exchange {args.exchange} needs symmetric memory"
                             + (" with a multicast mapping" if args.exchange == "nvls" else "") + ", not available here")
        else:
            exchange = args.exchange
    # SMs left free by the persistent row kernels for the exchange kernels that run under them (overlap mode):
    # NCCL needs its 16 CTAs (4 / 8 / 16 / 24 / 32 measured at N=8); the library's own kernel is as fast with 8
    # (N=4, c5: 911 G tok*V/s with 8 reserved SMs vs 887 with 16, profiles/r02k_xgpu_exchange.md)
    reserve_sms = nccl_ctas if exchange == "nccl" else int(os.environ.get("FDDM_XGPU_CTAS", "8"))
    overlap = world > 1 and (args.collectives == "overlap" or
                             (args.collectives == "auto" and (exchange != "nccl" or rows_per_gpu <= 40960)))
    def shard_checks(collective):
        res = shard_check(fb, dev, group, world, rank, per_rank_b=32, collective=collective)   # the path the timed step takes
        if res["ok"]:
            small = shard_check(fb, dev, group, world, rank, per_rank_b=8, collective=collective)   # and the small-batch path
            res = {**res, "small_batch_path": small, "ok": small["ok"]}
        res["exchange"] = collective
        return res

    shard = None
    if world > 1:
        shard = shard_checks(exchange)
        if not shard["ok"] and args.exchange == "auto" and exchange != "nccl":
            if rank == 0:
                print(f"bench.py: shard_check failed with --exchange {exchange} ({shard}); falling back to NCCL",
                      file=sys.stderr, flush=True)
            failed, exchange = shard, "nccl"
            reserve_sms = nccl_ctas
            overlap = args.collectives == "overlap" or (args.collectives == "auto" and rows_per_gpu <= 40960)
            shard = {**shard_checks(exchange), "own_kernel_exchange_failed": failed}
        if not shard["ok"]:
            if rank == 0:
                print(json.dumps({"error": "shard_check failed: the batch-sharded path disagrees with the whole-batch "
                                           "evaluation", "shard_check": shard, "n_gpus": world}), flush=True)
            teardown(world, dev, code=3)

    B, L, V, D, kind, sharded = shape_of(args.workload, world)
    Bg = B * world
    dtype = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[args.dtype]
    s_bytes = 4 if args.dtype == "f32" else 2
    d = synth(B, L, V, D, 1337 + rank, dtype, device=dev)
    # the same batch in pinned host memory, double-buffered on the device, for the end-to-end leg
    host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in d.items()}
    for k, v in d.items():
        host[k].copy_(v)
    torch.cuda.synchronize(dev)
    feats = [k for k in ("logits", "za", "zb") if k in d]
    for k in feats:
        d[k].requires_grad_(kind == "train")

    if overlap:
        fb.set_sm_reserve(reserve_sms)
        if exchange != "nccl":
            fb.losses._SymmExchange.of(group).max_ctas = reserve_sms
    sch = fb.DiscreteDiffusionScheduler(K=V, T=T_TRAIN, device=dev)
    # multi-GPU: the KL scalar is folded into the one scalar all-reduce at the end of the step
    ad = fb.SchedulerAdapter(sch, group=group, defer_reduce=(D > 0))
    dec = ResidentDecoder()
    smp = fb.DiffusionJumpySampler(sch, dec, K=V, T_train=T_TRAIN, T_infer=T_INFER, r=R_JUMP,
                                   greedy=(args.greedy if kind == "sampler" else False),
                                   sampling_mode=(args.sampling_mode if kind == "sampler" else "exact"), device=dev)
    cond = torch.zeros(B, 1, 1, device=dev)
    # device-side Philox {seed, offset}: read by the kernels, advanced by (captured) adds, so that a
    # CUDA-graph replay draws fresh noise exactly like an eager step would
    pstate = torch.tensor([1337 + rank, 0], dtype=torch.int64, device=dev)
    smp.philox_state = pstate
    x_init = torch.randint(0, V, (B, L), device=dev)

    def step(dd):
        if kind == "sampler":
            dec.logits = dd["logits"]
            ids, p_last = smp.sample(cond, L, x_init=x_init)
            return p_last.view(-1)[0].float(), ids
        for k in feats:
            dd[k].grad = None
        lfd_op = None
        if D > 0:
            lfd_op = fb.LfdPipeline(dd["za"], dd["zb"], LAMBDA, group=group, overlap=overlap, collective=exchange)
            lfd_op.stats()                                        # + all-reduce of the batch statistics
        pstate[1:].add_(8)
        xt = ad.sample_q(dd["x0"], dd["t"], philox_state=pstate)
        kl = ad.kl_term(xt, dd["x0"], dd["logits"], dd["t"], dd["mask"])
        if lfd_op is not None:
            lfd_op.xcov(piggyback=kl if world > 1 else None)      # + all-reduce of the covariance (carrying the KL partial sum)
        dec.logits = dd["logits"].detach()
        x_new, _ = smp._jump_once(dd["x0"], T_INFER, R_JUMP, cond, L, want_p=False)
        total = kl
        if lfd_op is not None:
            lfd = lfd_op.loss()
            total = kl + TAU * ad.w_t(dd["t"]).mean() * lfd.float()
        total.backward()
        if world > 1 and lfd_op is not None:                      # reported loss: global KL rode along with the covariance
            total = lfd_op.piggyback + (total.detach() - kl.detach())
        return total.detach(), x_new

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
        clocks.start()
    for _ in range(args.warmup):
        step(d)
    sync_all()
    # The step is ~30 short launches, so the whole step (forward, collectives, backward) is captured once in a
    # CUDA graph and replayed: every library call is stream-ordered and sync-free by contract.
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
                g_total, g_xnew = step(d)
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
    c_lo = clocks.mark() if rank == 0 else 0
    sync_all()
    ev0.record()
    h0 = time.perf_counter()
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            total, x_new = step(d)
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps       # CPU time to enqueue one step (no sync)
    ev1.record()
    sync_all()
    if graph is not None:
        total, x_new = g_total, g_xnew
        launches = launches_per_step * args.steps
    else:
        launches = fb._lib.launch_count() - n0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    per_step_elems = (4 if kind == "sampler" else 1) * B * L * V  # per GPU per step
    ms_step = ms_total / args.steps
    value = world * per_step_elems / (ms_step * 1e-3) / 1e9
    loss_val = float(total.detach())
    # the K timed steps can be over in milliseconds: keep the GPU under the SAME load for another ~0.6 s so that
    # the 50 ms clock sampler has samples taken under this load (the reported times are the K steps above only)
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        for _ in range(4):
            if graph is not None:
                graph.replay()
            else:
                step(d)
        torch.cuda.synchronize(dev)
    c_hi = clocks.mark() if rank == 0 else 0

    # ---- per-kernel roofline: the library's own CUDA events around every launch, eager steps --------
    prof_steps = max(2, min(args.steps, 5))
    fb._lib.profile_enable(True)
    for _ in range(prof_steps):
        step(d)
    torch.cuda.synchronize(dev)
    prof = fb._lib.profile_read()
    fb._lib.profile_enable(False)
    sync_all()

    # ---- end to end through the host API with HOST buffers (`e2e`) ------------------------------
    # two device buffer sets: the H->D copy of step i+1 (copy stream) overlaps the compute of step i
    dh = [{k: torch.empty_like(v, device=dev) for k, v in host.items()} for _ in range(2)]
    for buf in dh:
        for k in feats:
            buf[k].requires_grad_(kind == "train")
    res_host = torch.empty(1, dtype=torch.float32).pin_memory()
    ids_host = torch.empty(B, L, dtype=torch.int64).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = res_host.numel() * 4 + ids_host.numel() * 8
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def h2d_async(i):
        with torch.cuda.stream(copy_stream), torch.no_grad():
            copy_stream.wait_event(consumed[i])                   # the step that last read this buffer set is done
            for k, v in host.items():
                dh[i][k].copy_(v, non_blocking=True)
            copied[i].record(copy_stream)

    def e2e_run(n):
        cur = torch.cuda.current_stream(dev)
        for ev in consumed:
            ev.record(cur)
        h2d_async(0)
        for i in range(n):
            b = i & 1
            if i + 1 < n:
                h2d_async(b ^ 1)
            cur.wait_event(copied[b])
            tot, x_new = step(dh[b])
            consumed[b].record(cur)
            res_host.copy_(tot.detach().reshape(1), non_blocking=True)
            ids_host.copy_(x_new, non_blocking=True)
            cur.synchronize()                                     # the training loop's .item()

    e2e_run(2)
    sync_all()
    e2e_steps = max(3, min(args.steps, 10))
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    e2e_run(e2e_steps)
    ee1.record()
    sync_all()
    clk = clocks.stop(c_lo, c_hi) if rank == 0 else {}
    e2e_ms = max_over_ranks(ee0.elapsed_time(ee1)) / e2e_steps
    e2e_val = world * per_step_elems / (e2e_ms * 1e-3) / 1e9

    # ---- roofline ----------------------------------------------------------------------------------
    pk = peaks()
    valid_rows = int(host["mask"].sum()) if kind == "train" else B * L
    kernels, lfd_ms, prof_ms = kernel_roofline(prof, prof_steps, (B, L, V, D), args.dtype, valid_rows, pk,
                                               4 if kind == "sampler" else 0)
    rows = B * L
    if kind == "sampler":
        step_bytes = float(s_bytes) * V * rows * 5                # 4 reads of the logits + one write of p_x0
    else:
        step_bytes = float(s_bytes) * V * (valid_rows + rows) + float(s_bytes) * V * rows + 4.0 * rows * D * s_bytes
    traffic_db = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic_db = json.load(f)
    except Exception:
        pass
    for e in kernels:
        t = traffic_db.get(f"{e['kernel']}:{args.dtype}:V{V}") or traffic_db.get(f"{e['kernel']}:{args.dtype}:D{D}")
        if t:                                                     # ncu --set full capture, scaled by rows to this launch
            e["traffic"] = round(t["bytes_per_row"] * rows, 0); e["traffic_source"] = t["source"]
    dom = next((e for e in kernels if "frac" in e), None)         # largest time share with a defined roofline
    roofline = {
        "bound": dom["bound"] if dom else "hbm", "kernel": dom["kernel"] if dom else None,
        "achieved": dom["achieved"] if dom else None, "peak": dom["peak"] if dom else pk["hbm_gbs"],
        "unit": dom["unit"] if dom else "GB/s", "frac": dom["frac"] if dom else None,
        "traffic": dom.get("traffic") if dom else None, "traffic_source": dom.get("traffic_source") if dom else None,
        "selected_by": "largest share of the step's device time among the library's kernels",
        "timed_by": f"CUDA events recorded by the library around every launch, {prof_steps} eager steps",
        "peak_source": pk["source"],
        "kernels": kernels,
        "lfd_chain": None if D == 0 else {
            "ms_per_step": round(lfd_ms, 4), "algorithmic_bytes": 4.0 * rows * D * s_bytes, "useful_flops": 6.0 * rows * D * D,
            "frac_hbm": round(4.0 * rows * D * s_bytes / (lfd_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
            "frac_tensor": round(6.0 * rows * D * D / (lfd_ms * 1e-3) / 1e12 / pk["bf16_tflops"], 4)},
        "step_algorithmic_bytes": step_bytes,
        "step_frac": round(step_bytes / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
        "step_frac_of_nominal_8TBs": round(step_bytes / (ms_step * 1e-3) / 1e9 / 8000.0, 4),
        "kernel_ms_sum_eager": round(prof_ms, 4),
    }

    cpu = eager = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            ev, ems, esample, ekind = run_baseline(args.workload, 3, 1, dev, EAGER_SLICE_B, args.sampling_mode, args.greedy)
            eager = {"value": round(ev, 3), "unit": UNIT, "kind": ekind, "device": torch.cuda.get_device_name(dev),
                     "ms_per_step_slice": round(ems, 2), "sample": esample}
        except Exception as e:                                    # e.g. out of memory: report, never fail the line
            eager = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        v, ms, sample, ckind = run_baseline(args.workload, 3, 1, "cpu", CPU_SLICE_B, args.sampling_mode, args.greedy)
        cpu = {"value": round(v, 5), "unit": UNIT, "cores": cores, "kind": ckind, "sample": sample,
               "ms_per_step_slice": round(ms, 1), "cpu_model": cpu_model()}

    if rank == 0:
        what = ("4 fused jumps of DiffusionJumpySampler.sample()" if kind == "sampler" else
                "sample_q + KL fwd+bwd + exact jump/resample" + (" + L_fd fwd+bwd" if D > 0 else ""))
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic (seeded N(0,1) logits, padded ids, correlated features)",
            "config": {"workload": f"{args.workload}: global B={Bg} L={L} V={V} D={D}, B={B} per GPU"
                                   + (" (BASELINE configs[4], batch-sharded)" if args.workload == "c5" else "")
                                   + f"; step = {what}",
                       "global_batch": Bg, "seq_len": L, "vocab": V, "d_proj": D, "T": T_TRAIN,
                       "parallelism": f"batch-sharded x{world}" + ((" (NCCL all-reduce: KL scalar, L_fd stats/cov/bn sums" if exchange == "nccl" else
                                                                      f" (the library's own {exchange} all-reduce kernel on symmetric memory: L_fd stats / cov + KL scalar / bn sums")
                                                                     + (f"; the two forward all-reduces overlap the KL / jump kernels, {reserve_sms} SMs reserved" if overlap else "")
                                                                     + ")" if world > 1 else ""),
                       "sampling_mode": smp.sampling_mode, "greedy": smp.greedy,
                       "l2": f"inputs {h2d / 1e6:.0f} MB/step per GPU > 126 MB L2 (no flush needed)"},
            "e2e": {"value": round(e2e_val, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_ms, 3), "steps": e2e_steps,
                    "how": "pinned host buffers, H->D of step i+1 on a copy stream overlapping step i, loss + ids read back every step"},
            "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_ms, 4),
            "cuda_graph": graph is not None,
            "clocks": clk,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "eager_b200": eager,
            "loss": loss_val,
            "shard_check": shard,
        }
        print(json.dumps(line), flush=True)
    # the captured graph holds NCCL kernels (a reference on the communicator): it must die before the group does
    graph = g_total = g_xnew = total = x_new = None
    teardown(world, dev)


def teardown(world, dev, code=0):
    """Leave the process without hanging.  Round-1 incident: every multi-rank run hung at
    exit once the step was a CUDA graph, because destroy_process_group() was called while the captured graph --
    which holds NCCL kernels and so a reference on the communicator -- was still alive.  Order here: the JSON
    line is already printed and flushed; the caller has dropped the graph and everything captured with it;
    synchronise, barrier, destroy the group under a watchdog, and finally _exit so that no destructor can block."""
    import gc
    import threading
    sys.stdout.flush(); sys.stderr.flush()
    if world <= 1:
        if code:
            sys.exit(code)
        return
    wd = threading.Timer(45.0, lambda: os._exit(code))         # the result is out: never hold the GPUs hostage
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
    """`--impl reference`: the reference path's CPU implementation on the host cores -- the unmodified
    reference when it is importable ($FDDM_REF, /root/reference, baseline/_ref), else the torch-eager port --
    same metric/unit/config, each step a bounded batch slice.  Under torchrun only rank 0 works."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    B, L, V, D, kind, sharded = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    v, ms, sample, ckind = run_baseline(args.workload, args.steps, args.warmup, "cpu", CPU_SLICE_B, args.sampling_mode,
                                        args.greedy)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 5), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (same generator as the B200 arm)",
        "config": {"workload": f"{args.workload}: global B={B} L={L} V={V} D={D}; each step is a bounded sample "
                               f"(batch slice B={min(CPU_SLICE_B, B)}) of it",
                   "seq_len": L, "vocab": V, "d_proj": D, "T": T_TRAIN, "parallelism": f"host CPU, {cores} threads"},
        "cpu_baseline": {"value": round(v, 5), "unit": UNIT, "cores": cores, "kind": ckind, "sample": sample,
                         "cpu_model": cpu_model()},
        "e2e": {"value": round(v, 5), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"])
    ap.add_argument("--sampling-mode", default="exact", choices=["exact", "fast"], help="c3 only")
    ap.add_argument("--greedy", action="store_true", help="c3 only: argmax instead of Categorical")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "p2p", "nvls"],
                    help="N>1: how the L_fd exchange buffers are summed: ncclAllReduce, or the library's own kernel on "
                         "symmetric memory (p2p: peer loads/stores; nvls: in-switch multicast reduction)")
    ap.add_argument("--collectives", default="auto", choices=["auto", "overlap", "serial"],
                    help="N>1: run L_fd's forward all-reduces under the KL / jump kernels (side stream) or in order")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and eager_b200 legs")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
