#!/usr/bin/env python
"""Generate golden input/output vectors by running the UNMODIFIED reference implementation.

Run in the build container only (the reference cannot travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [--ref /root/reference]

It imports `fddm.sched.diffusion_scheduler`, `losses.fddm_losses`, `sampler.jumpy_sampler` and
`train.SchedulerAdapter` from the reference tree, feeds them seeded numpy inputs on CPU and
stores inputs + outputs as small compressed fixtures in tests/golden/*.npz.  Randomness inside
the reference (`torch.multinomial`, `Categorical.sample`) is replaced by *injected*
exponential noise through the documented identity multinomial(p,1) == argmax(p / E), which this
script first proves bit-exact against torch's own sampler on CPU (same generator state).
Nothing from the reference is copied into the repo; only numbers are.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("FDDM_REF", "/root/reference"))
    args = ap.parse_args()
    sys.dont_write_bytecode = True
    sys.path.insert(0, args.ref)
    import warnings
    warnings.filterwarnings("ignore")
    import torch
    torch.set_num_threads(4)
    from fddm.sched.diffusion_scheduler import DiscreteDiffusionScheduler
    from losses.fddm_losses import lfd_loss
    from sampler.jumpy_sampler import DiffusionJumpySampler
    import train as ref_train

    dev = torch.device("cpu")
    out = {}

    # ---------------------------------------------------------------- injected-noise shim
    class Injector:
        """Replaces torch.multinomial(p, 1) by argmax(p / E) with E popped from a queue."""
        def __init__(self):
            self.queue = []
            self.orig = torch.multinomial

        def __call__(self, probs, num_samples, replacement=False, *, generator=None):
            assert num_samples == 1
            E = self.queue.pop(0)
            E = torch.as_tensor(E, dtype=probs.dtype).reshape(probs.shape)
            return (probs / E).argmax(dim=-1, keepdim=True)

    # prove the identity on CPU: same generator state -> same ids
    g = torch.Generator().manual_seed(123)
    p = torch.rand(64, 257, generator=g) + 1e-3
    g1 = torch.Generator().manual_seed(7)
    ids_torch = torch.multinomial(p, 1, generator=g1)
    g2 = torch.Generator().manual_seed(7)
    E = torch.empty_like(p).exponential_(1, generator=g2)
    ids_race = (p / E).argmax(-1, keepdim=True)
    assert torch.equal(ids_torch, ids_race), "multinomial(p,1) != argmax(p/E) on this torch build"
    out["identity_checked_torch_version"] = np.array(torch.__version__)

    inj = Injector()

    rng = np.random.default_rng(1337)

    def exp_noise(shape):
        u = rng.random(shape, dtype=np.float64)
        u = np.clip(u, 1e-12, 1.0 - 1e-12)
        return (-np.log(u)).astype(np.float32)

    # ---------------------------------------------------------------- tables
    sch = DiscreteDiffusionScheduler(K=8000, T=200, device=dev, beta_max=0.2)
    out["tab_betas"] = sch.betas.numpy().copy()
    out["tab_alpha_bar"] = sch.alpha_bar.numpy().copy()
    sch_small = DiscreteDiffusionScheduler(K=53, T=50, device=dev, beta_max=0.3)
    out["tab53_betas"] = sch_small.betas.numpy().copy()
    out["tab53_alpha_bar"] = sch_small.alpha_bar.numpy().copy()

    # ---------------------------------------------------------------- q_sample (a2)
    # K=8000 rows with t in {1,2,200} exercise the eps clamp (quirk Q4)
    t_q = np.array([1, 2, 200, 77], dtype=np.int64)
    x0_ids = rng.integers(0, 8000, size=(4, 2))
    oh = torch.zeros(4, 2, 8000)
    oh.scatter_(-1, torch.from_numpy(x0_ids).unsqueeze(-1), 1.0)
    out["qs_t"] = t_q
    out["qs_x0_ids"] = x0_ids
    out["qs_out"] = sch.q_sample(oh, torch.from_numpy(t_q)).numpy()
    # soft (non one-hot) input on the small scheduler
    x0_soft = rng.dirichlet(np.ones(53) * 0.3, size=(5, 7)).astype(np.float32)
    t_s = rng.integers(1, 51, size=5).astype(np.int64)
    t_s[0] = 1; t_s[1] = 50
    out["qs53_x0"] = x0_soft
    out["qs53_t"] = t_s
    out["qs53_out"] = sch_small.q_sample(torch.from_numpy(x0_soft), torch.from_numpy(t_s)).numpy()

    # ---------------------------------------------------------------- sample_q (a3)
    adapter = ref_train.SchedulerAdapter(sch)
    B, L = 6, 5
    x0 = rng.integers(0, 8000, size=(B, L)).astype(np.int64)
    t = np.array([1, 2, 3, 60, 130, 200], dtype=np.int64)
    E = exp_noise((B * L, 8000))
    inj.queue.append(E)
    torch.multinomial = inj
    try:
        xt = adapter.sample_q(torch.from_numpy(x0), torch.from_numpy(t)).numpy()
    finally:
        torch.multinomial = inj.orig
    out["sq_x0"] = x0; out["sq_t"] = t; out["sq_E_seed"] = np.array(0)
    # the noise is 960 KB raw; keep only what decides each row: store it in float32 compressed
    out["sq_E"] = E
    out["sq_xt"] = xt
    # small-K variant with many corrupted tokens
    adapter_s = ref_train.SchedulerAdapter(sch_small)
    B, L = 8, 16
    x0s = rng.integers(0, 53, size=(B, L)).astype(np.int64)
    ts = rng.integers(1, 51, size=B).astype(np.int64)
    Es = exp_noise((B * L, 53))
    inj.queue.append(Es)
    torch.multinomial = inj
    try:
        xts = adapter_s.sample_q(torch.from_numpy(x0s), torch.from_numpy(ts)).numpy()
    finally:
        torch.multinomial = inj.orig
    out["sq53_x0"] = x0s; out["sq53_t"] = ts; out["sq53_E"] = Es; out["sq53_xt"] = xts

    # ---------------------------------------------------------------- q_posterior (a4)
    B, L, K = 5, 4, 53
    xt_p = rng.dirichlet(np.ones(K) * 0.5, size=(B, L)).astype(np.float32)
    xh_p = rng.dirichlet(np.ones(K) * 0.2, size=(B, L)).astype(np.float32)
    # make two rows one-hot (the sampler's use)
    xt_p[0, 0] = 0; xt_p[0, 0, 3] = 1
    xt_p[1, 2] = 0; xt_p[1, 2, 52] = 1
    t_p = np.array([1, 2, 17, 50, 33], dtype=np.int64)
    out["qp_xt"] = xt_p; out["qp_xh"] = xh_p; out["qp_t"] = t_p
    out["qp_out"] = sch_small.q_posterior(torch.from_numpy(xt_p), torch.from_numpy(xh_p),
                                          torch.from_numpy(t_p)).numpy()

    # ---------------------------------------------------------------- multi-step (a5)
    cases = []
    for ci, (tvec, delta) in enumerate([
        (np.array([50, 50, 50, 50, 50]), 5),
        (np.array([5, 5, 5, 5, 5]), 5),            # t_target == 0
        (np.array([9, 20, 50, 3, 14]), 7),         # delta clipped to t.min()=3 (Q10)
        (np.array([2, 2, 2, 2, 2]), 1),
        (np.array([1, 7, 9, 4, 4]), 3),            # clipped to 1
        (np.array([10, 10, 10, 10, 10]), 0),       # delta<=0 -> identity
    ]):
        tv = tvec.astype(np.int64)
        res = sch_small.q_posterior_multi_step(torch.from_numpy(xt_p), torch.from_numpy(xh_p),
                                               torch.from_numpy(tv), int(delta)).numpy()
        out[f"ms{ci}_t"] = tv; out[f"ms{ci}_delta"] = np.array(delta); out[f"ms{ci}_out"] = res
        cases.append(ci)
    out["ms_n"] = np.array(len(cases))
    # coefficients at the production table (K=8000, T=200): recover (a_cum, b_cum) by probing the
    # reference with one-hot rows -- A_k = a*[k=xt] + b ; here we instead re-run its loop output
    # via two probes: x_t one-hot, x0hat uniform => posterior closed form is not needed; we only
    # store full outputs for a tiny L.
    K2 = 8000
    xt2 = np.zeros((3, 1, K2), dtype=np.float32); xt2[:, 0, 11] = 1
    xh2 = rng.dirichlet(np.ones(K2) * 0.05, size=(3, 1)).astype(np.float32)
    for ci, (tv, delta) in enumerate([(np.array([200, 200, 200]), 5), (np.array([20, 20, 20]), 5),
                                      (np.array([5, 5, 5]), 5)]):
        res = sch.q_posterior_multi_step(torch.from_numpy(xt2), torch.from_numpy(xh2),
                                         torch.from_numpy(tv.astype(np.int64)), int(delta)).numpy()
        out[f"ms8k{ci}_t"] = tv.astype(np.int64); out[f"ms8k{ci}_delta"] = np.array(delta)
        out[f"ms8k{ci}_out"] = res
    out["ms8k_xt_id"] = np.array(11); out["ms8k_xh"] = xh2

    # ---------------------------------------------------------------- kl_term (a7)
    def run_kl(tag, B, L, V, sched, tt, dtype, masked, sigma):
        ad = ref_train.SchedulerAdapter(sched)
        x0 = rng.integers(0, V, size=(B, L)).astype(np.int64)
        xt = x0.copy()
        flip = rng.random((B, L)) < 0.5
        xt[flip] = rng.integers(0, V, size=int(flip.sum()))
        z = (rng.standard_normal((B, L, V)) * sigma).astype(np.float32)
        # make the model mostly right on some rows so both KL regimes appear
        for b in range(B):
            for l in range(0, L, 2):
                z[b, l, x0[b, l]] += 6.0
        zt = torch.from_numpy(z).to(dtype).requires_grad_(True)
        mask = None
        if masked:
            lens = rng.integers(0, L + 1, size=B)
            lens[0] = L
            if B > 1:
                lens[1] = 0                                   # an empty sample (Q6)
            mask = (np.arange(L)[None, :] < lens[:, None])
        loss = ad.kl_term(torch.from_numpy(xt), torch.from_numpy(x0), zt, torch.from_numpy(tt),
                          None if mask is None else torch.from_numpy(mask))
        loss.backward()
        out[f"kl_{tag}_x0"] = x0; out[f"kl_{tag}_xt"] = xt; out[f"kl_{tag}_t"] = tt
        out[f"kl_{tag}_logits"] = zt.detach().float().numpy()
        out[f"kl_{tag}_mask"] = np.array([]) if mask is None else mask
        out[f"kl_{tag}_loss"] = loss.detach().float().numpy()
        out[f"kl_{tag}_loss_dtype"] = np.array(str(loss.dtype))
        out[f"kl_{tag}_grad"] = zt.grad.float().numpy()
        out[f"kl_{tag}_grad_dtype"] = np.array(str(zt.grad.dtype))

    run_kl("f32", 6, 5, 53, sch_small, np.array([1, 2, 50, 17, 33, 5], dtype=np.int64), torch.float32, False, 1.0)
    run_kl("f32m", 6, 5, 53, sch_small, np.array([1, 2, 50, 17, 33, 5], dtype=np.int64), torch.float32, True, 3.0)
    run_kl("bf16m", 6, 5, 53, sch_small, np.array([1, 2, 50, 17, 33, 5], dtype=np.int64), torch.bfloat16, True, 2.0)
    run_kl("f16", 6, 5, 53, sch_small, np.array([1, 2, 50, 17, 33, 5], dtype=np.int64), torch.float16, False, 2.0)
    run_kl("f32k8", 4, 3, 8000, sch, np.array([1, 2, 200, 100], dtype=np.int64), torch.float32, True, 2.0)

    # ---------------------------------------------------------------- lfd_loss (a8)
    def run_lfd(tag, B, T, D, rho, lam, dtype):
        za = rng.standard_normal((B, T, D)).astype(np.float32)
        zb = (rho * za + np.sqrt(1 - rho * rho) * rng.standard_normal((B, T, D))).astype(np.float32)
        ta = torch.from_numpy(za).to(dtype).requires_grad_(True)
        tb = torch.from_numpy(zb).to(dtype).requires_grad_(True)
        loss = lfd_loss(ta, tb, lambda_offdiag=lam)
        loss.backward()
        out[f"lfd_{tag}_za"] = ta.detach().float().numpy(); out[f"lfd_{tag}_zb"] = tb.detach().float().numpy()
        out[f"lfd_{tag}_lam"] = np.array(lam)
        out[f"lfd_{tag}_loss"] = loss.detach().float().numpy()
        out[f"lfd_{tag}_loss_dtype"] = np.array(str(loss.dtype))
        out[f"lfd_{tag}_ga"] = ta.grad.float().numpy(); out[f"lfd_{tag}_gb"] = tb.grad.float().numpy()

    run_lfd("r0", 8, 5, 16, 0.0, 5e-3, torch.float32)
    run_lfd("r9", 8, 5, 16, 0.9, 5e-3, torch.float32)
    run_lfd("l1", 6, 3, 24, 0.5, 1.0, torch.float32)
    run_lfd("bf16", 8, 5, 16, 0.9, 5e-3, torch.bfloat16)

    # ---------------------------------------------------------------- jumpy sampler (a9/a10)
    class TableDecoder:
        """Dummy decoder: returns pre-generated logits per call (the decoder is out of scope)."""
        def __init__(self, tables, dtype):
            self.tables = tables; self.calls = 0; self.dtype = dtype; self.seen = []

        def __call__(self, x_t_idx, t, cond_c):
            self.seen.append((x_t_idx.numpy().copy(), t.numpy().copy()))
            z = torch.from_numpy(self.tables[self.calls]).to(self.dtype)
            self.calls += 1
            return z

    def run_sampler(tag, sched, K, T_infer, r, B, L, mode, pmode, greedy, temp, dtype, sigma):
        n_jumps = -(-T_infer // r)
        tables = [(rng.standard_normal((B, L, K)) * sigma).astype(np.float32) for _ in range(n_jumps)]
        dec = TableDecoder(tables, dtype)
        smp = DiffusionJumpySampler(sched, dec, K=K, T_train=sched.T, T_infer=T_infer, r=r,
                                    greedy=greedy, posterior_mode=pmode, sampling_mode=mode,
                                    temperature=temp, device=dev)
        x_T = rng.integers(0, K, size=(B, L)).astype(np.int64)
        noises = []
        needs_noise = (pmode != "max") and (not greedy)
        if needs_noise:
            for _ in range(n_jumps):
                En = exp_noise((B * L, K)); noises.append(En); inj.queue.append(En)
        orig_randint = torch.randint
        torch.randint = lambda *a, **k: torch.from_numpy(x_T.copy())
        torch.multinomial = inj
        try:
            x0_idx, p_last = smp.sample(torch.zeros(B, 3, 4), L)
        finally:
            torch.randint = orig_randint
            torch.multinomial = inj.orig
        out[f"js_{tag}_logits"] = np.stack([torch.from_numpy(tb).to(dtype).float().numpy() for tb in tables])
        out[f"js_{tag}_xT"] = x_T
        out[f"js_{tag}_noise"] = np.stack(noises) if noises else np.array([])
        out[f"js_{tag}_x_seen"] = np.stack([s[0] for s in dec.seen])     # x_t fed to the decoder per jump
        out[f"js_{tag}_t_seen"] = np.stack([s[1] for s in dec.seen])
        out[f"js_{tag}_x0"] = x0_idx.numpy()
        out[f"js_{tag}_p_last"] = p_last.float().numpy()
        out[f"js_{tag}_p_last_dtype"] = np.array(str(p_last.dtype))
        out[f"js_{tag}_cfg"] = np.array([K, sched.T, T_infer, r, B, L, int(greedy)], dtype=np.int64)
        out[f"js_{tag}_mode"] = np.array([mode, pmode, str(temp), str(dtype)])

    run_sampler("ex_g", sch_small, 53, 20, 5, 3, 6, "exact", "average", True, 1.0, torch.float32, 2.0)
    run_sampler("ex_s", sch_small, 53, 20, 5, 3, 6, "exact", "average", False, 1.0, torch.float32, 2.0)
    run_sampler("ex_map", sch_small, 53, 7, 3, 3, 6, "exact", "map", False, 1.0, torch.float32, 2.0)
    run_sampler("ex_max", sch_small, 53, 7, 3, 3, 6, "exact", "max", False, 1.0, torch.float32, 2.0)
    run_sampler("fa_g", sch_small, 53, 20, 5, 3, 6, "fast", "average", True, 1.0, torch.float32, 2.0)
    run_sampler("fa_s", sch_small, 53, 20, 5, 3, 6, "fast", "average", False, 1.0, torch.float32, 2.0)
    run_sampler("fa_st", sch_small, 53, 20, 5, 3, 6, "fast", "average", False, 0.7, torch.float32, 2.0)
    run_sampler("ex_st", sch_small, 53, 20, 5, 3, 6, "exact", "average", False, 1.3, torch.float32, 2.0)
    run_sampler("ex_s_bf16", sch_small, 53, 20, 5, 3, 6, "exact", "average", False, 1.0, torch.bfloat16, 2.0)
    run_sampler("fa_g_bf16", sch_small, 53, 20, 5, 3, 6, "fast", "average", True, 1.0, torch.bfloat16, 2.0)
    run_sampler("ex_s_8k", sch, 8000, 20, 5, 2, 2, "exact", "average", False, 1.0, torch.float32, 3.0)

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")


if __name__ == "__main__":
    main()
