#!/usr/bin/env python
"""One small invocation of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool <tool> python scripts/sanitize_smoke.py
Shapes are tiny but cover the TMA-ring row kernels (register-resident and streamed), the mbarrier/TMEM
contraction kernels (split-K forward, persistent 2-CTA-cluster backward) and both L_fd row orders."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fddm-asr_b200")):
    sys.path.insert(0, p)
import torch
import fddm_b200 as fb

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
T = 50
for V, dt in ((4000, torch.float32), (8000, torch.bfloat16), (1001, torch.float32)):
    B, L = 2, 24
    sch = fb.DiscreteDiffusionScheduler(K=V, T=T, device=dev)
    ad = fb.SchedulerAdapter(sch)
    logits = (torch.randn(B, L, V, generator=g, device=dev) * 2).to(dt).requires_grad_(True)
    x0 = torch.randint(0, V, (B, L), generator=g, device=dev)
    t = torch.randint(1, T + 1, (B,), generator=g, device=dev)
    mask = torch.rand(B, L, generator=g, device=dev) < 0.7
    xt = ad.sample_q(x0, t)
    E = torch.empty(B, L, V, device=dev).exponential_(generator=g)
    xt2 = ad.sample_q(x0, t, exp_noise=E)
    for _ in range(2):                                   # twice: the self-resetting work counters are reused
        logits.grad = None
        ad.kl_term(xt, x0, logits, t, mask).backward()
    with torch.no_grad():
        ad.kl_term(xt, x0, logits.detach(), t, None)
    for mode in ("exact", "fast"):
        for greedy in (True, False):
            smp = fb.DiffusionJumpySampler(sch, lambda x, tt, c: logits.detach(), K=V, T_train=T, T_infer=20, r=5,
                                           greedy=greedy, sampling_mode=mode, device=dev)
            smp.sample(torch.zeros(B, 1, 1, device=dev), L)
    smp.noise_fn = lambda step, shape: E
    smp._jump_once(x0, 10, 5, torch.zeros(B, 1, 1, device=dev), L)
    if V == 4000:
        p = torch.softmax(logits.detach().float(), -1)
        q = sch.q_sample(p, t); sch.q_posterior(q, p, t); sch.q_posterior_multi_step(q, p, t, 3)
for B, Tn, D in ((4, 6, 16), (40, 5, 264)):               # natural row order / tb-major with batch padding
    za = torch.randn(B, Tn, D, generator=g, device=dev, requires_grad=True)
    zb = (0.8 * za.detach() + 0.6 * torch.randn(B, Tn, D, generator=g, device=dev)).requires_grad_(True)
    for _ in range(2):
        za.grad = zb.grad = None
        fb.lfd_loss(za, zb, 5e-3).backward()
torch.cuda.synchronize()
print("sanitize_smoke ok, launches:", fb._lib.launch_count())
