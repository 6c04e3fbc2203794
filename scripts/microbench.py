#!/usr/bin/env python
"""Kernel micro-benchmarks on one GPU (CUDA events around K back-to-back launches, inputs >> L2):
    python scripts/microbench.py [jump] [kl] [lfd] [--B 64] [--L 256] [--V 8000] [--D 768] [--dtype f32] [--iters 20]
Prints one line per case: ms per launch, algorithmic GB/s (or TFLOP/s) and the fraction of the measured peak."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fddm-asr_b200")):
    sys.path.insert(0, p)
import torch
import fddm_b200 as fb
import bench


KERNEL_MS = {}


def timed(fn, iters, warm=3):
    """ms per call between two CUDA events around `iters` back-to-back calls (includes host-side gaps when the
    host cannot enqueue as fast as the GPU executes); KERNEL_MS additionally gets the device time of the
    library's kernels alone, from the events the library records around each launch."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    fb._lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    prof = fb._lib.profile_read()
    fb._lib.profile_enable(False)
    KERNEL_MS.clear()
    KERNEL_MS.update({k: ms / n for k, (n, ms) in prof.items()})
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["jump", "kl", "lfd"])
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--L", type=int, default=256)
    ap.add_argument("--V", type=int, default=8000)
    ap.add_argument("--D", type=int, default=768)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    pk = bench.peaks()
    dt = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[a.dtype]
    s = 4 if a.dtype == "f32" else 2
    B, L, V, D = a.B, a.L, a.V, a.D
    d = bench.synth(B, L, V, D, 1337, dt, device=dev)
    rows = B * L
    sch = fb.DiscreteDiffusionScheduler(K=V, T=200, device=dev)

    def rec(name, ms, by=None, fl=None):
        e = {"case": name, "ms": round(ms, 4)}
        big = max(KERNEL_MS.items(), key=lambda kv: kv[1]) if KERNEL_MS else None
        if big and not name.startswith("lfd"):                    # single-kernel cases: report the kernel's own time
            e["kernel_ms"] = round(big[1], 4)
            ms = big[1]
        if by:
            e.update(GBps=round(by / ms / 1e6, 1), frac_hbm=round(by / ms / 1e6 / pk["hbm_gbs"], 4))
        if fl:
            e.update(TFLOPs=round(fl / ms / 1e9, 1), frac_tensor=round(fl / ms / 1e9 / pk["bf16_tflops"], 4))
        print(json.dumps(e), flush=True)

    if "jump" in a.what:
        class Dec:
            def __call__(self, x, t, c):
                return d["logits"]
        cond = torch.zeros(B, 1, 1, device=dev)
        ps = torch.tensor([1, 0], dtype=torch.int64, device=dev)
        for mode in ("exact", "fast"):
            for greedy in (False, True):
                smp = fb.DiffusionJumpySampler(sch, Dec(), K=V, T_train=200, T_infer=20, r=5, greedy=greedy,
                                               sampling_mode=mode, device=dev)
                smp.philox_state = ps
                ms = timed(lambda: smp._jump_once(d["x0"], 20, 5, cond, L, want_p=False), a.iters)
                rec(f"jump {mode} {'greedy' if greedy else 'categorical(philox)'} V={V} {a.dtype}", ms, s * V * rows)
        smp = fb.DiffusionJumpySampler(sch, Dec(), K=V, T_train=200, T_infer=20, r=5, greedy=False, device=dev)
        smp.philox_state = ps
        ms = timed(lambda: smp._jump(d["x0"], 5, 5, cond, L, want_p=True, want_argmax=True, step=3), a.iters)
        rec(f"jump exact categorical last (p_x0 + argmax written) V={V} {a.dtype}", ms, 2 * s * V * rows)
    if "kl" in a.what:
        ad = fb.SchedulerAdapter(sch)
        xt = ad.sample_q(d["x0"], d["t"])
        valid = int(d["mask"].sum())
        lg = d["logits"].detach().requires_grad_(True)
        ms = timed(lambda: ad.kl_term(xt, d["x0"], lg, d["t"], d["mask"]), a.iters)
        rec(f"kl fwd+bwd V={V} {a.dtype} valid={valid / rows:.2f}", ms, s * V * (valid + rows))
        lg0 = d["logits"].detach()
        with torch.no_grad():
            ms = timed(lambda: ad.kl_term(xt, d["x0"], lg0, d["t"], d["mask"]), a.iters)
        rec(f"kl fwd V={V} {a.dtype}", ms, s * V * valid)
        ms = timed(lambda: ad.kl_term(xt, d["x0"], lg, d["t"], None), a.iters)
        rec(f"kl fwd+bwd unmasked V={V} {a.dtype}", ms, 2 * s * V * rows)
    if "lfd" in a.what and D > 0:
        za = d["za"].detach().requires_grad_(True); zb = d["zb"].detach().requires_grad_(True)

        def fwdbwd():
            za.grad = None; zb.grad = None
            fb.lfd_loss(za, zb, 5e-3).backward()
        for _ in range(2):
            fwdbwd()
        ms = timed(fwdbwd, a.iters, warm=0)
        rec(f"lfd fwd+bwd B={B} T={L} D={D} {a.dtype}", ms, 4 * rows * D * s, 6.0 * rows * D * D)
        for k, t in sorted(KERNEL_MS.items(), key=lambda kv: -kv[1]):
            print(f"    {k:40s} {t * 1e3:9.1f} us")


if __name__ == "__main__":
    main()
