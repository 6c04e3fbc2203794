"""GPU tests at BASELINE.json's full sizes (configs c2, c3, c4 and the c5 per-GPU shard), where the numpy
oracle would take minutes.  The independent checker here is the torch restatement of the reference
(oracle/fddm_torch_port.py, pinned to the reference-generated golden vectors by the CPU tests) evaluated in
FP64 ON THE GPU over WHOLE tensors (test_kl_fullsize_vs_port_fp64, test_jumpy_sampler_c3_vs_port,
test_lfd_fullsize_properties), plus size-independent properties of the domain and consistency between the
fused kernels and this library's own dense-API kernels."""
import numpy as np
import pytest
import torch

from oracle import fddm_oracle as O
from oracle import fddm_torch_port as P

pytestmark = pytest.mark.gpu
T_TRAIN = 200


@pytest.fixture(scope="module")
def fb():
    import fddm_b200
    assert torch.cuda.is_available() and fddm_b200._lib.MISSING == []
    return fddm_b200


def sched(fb, K):
    return fb.DiscreteDiffusionScheduler(K=K, T=T_TRAIN, device=torch.device("cuda"))


def synth(B, L, V, dtype, seed=1337):
    g = torch.Generator(device="cuda").manual_seed(seed)
    logits = (torch.randn(B, L, V, generator=g, device="cuda") * 2.0).to(dtype)
    x0 = torch.randint(1, V, (B, L), generator=g, device="cuda")
    lens = torch.randint(int(0.4 * L), L + 1, (B,), generator=g, device="cuda")
    mask = torch.arange(L, device="cuda")[None, :] < lens[:, None]
    x0 = torch.where(mask, x0, torch.zeros_like(x0))
    t = torch.randint(1, T_TRAIN + 1, (B,), generator=g, device="cuda")
    t[0], t[1], t[-1] = 1, 2, T_TRAIN
    return logits, x0, mask, t


@pytest.mark.parametrize("B,L,V,dtype", [(32, 128, 8000, torch.float32),       # c2
                                         (64, 256, 8000, torch.float32),       # c5 per-GPU shard
                                         (64, 256, 32000, torch.float32),      # c4 fp32
                                         (64, 256, 32000, torch.bfloat16)])    # c4 bf16
def test_kl_fullsize_properties(fb, B, L, V, dtype):
    s = sched(fb, V)
    ad = fb.SchedulerAdapter(s)
    logits, x0, mask, t = synth(B, L, V, dtype)
    xt = ad.sample_q(x0, t)
    assert xt.shape == x0.shape and int(xt.min()) >= 0 and int(xt.max()) < V
    lg = logits.clone().requires_grad_(True)
    loss = ad.kl_term(xt, x0, lg, t, mask)
    loss.backward()
    g = lg.grad
    assert torch.isfinite(loss) and float(loss) > 0
    assert torch.isfinite(g).all()
    # softmax gradient property: every row of d loss / d logits sums to zero
    rs = g.float().sum(-1)
    scale = g.float().abs().sum(-1).clamp_min(1e-30)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert float((rs.abs() / scale).max()) < tol
    assert float(g[~mask].abs().max()) == 0.0                       # masked rows: exactly zero
    # linearity in the batch: the loss is the mean of per-sample terms -> halves average to the whole
    with torch.no_grad():
        h = B // 2
        l1 = ad.kl_term(xt[:h], x0[:h], logits[:h], t[:h], mask[:h])
        l2 = ad.kl_term(xt[h:], x0[h:], logits[h:], t[h:], mask[h:])
        whole = ad.kl_term(xt, x0, logits, t, mask)
    assert abs(float(whole) - 0.5 * (float(l1) + float(l2))) <= 2e-6 * abs(float(whole))
    assert abs(float(whole) - float(loss.detach())) <= 1e-5 * abs(float(whole))
    # oracle on a sample of rows (fp64), including the edge samples t=1, t=2, t=T
    rng = np.random.default_rng(0)
    bs = np.array([0, 1, B - 1, int(rng.integers(2, B - 1))])
    ls = np.sort(rng.choice(L, size=3, replace=False))
    sub = np.ix_(bs, ls)
    io = "f32" if dtype == torch.float32 else "bf16"
    tok, grad = O.kl_token_terms(xt.cpu().numpy()[sub], x0.cpu().numpy()[sub], logits.float().cpu().numpy()[sub],
                                 t.cpu().numpy()[bs], s.betas.cpu().numpy(), io_dtype=io, dtype=np.float64, want_grad=True)
    m = mask.cpu().numpy()
    w = (m / (m.sum(1, keepdims=True) + 1e-8) / B)[sub]
    got = g.float().cpu().numpy()[sub]
    want = grad * w[..., None]
    gtol = 2e-4 if dtype == torch.float32 else 2e-2               # fp32: includes the reference's own t<=2 noise
    assert np.abs(got - want).max() <= gtol * np.abs(want).max()


def port_kl(xt, x0, logits, t, betas, mask, dtype, chunk=8):
    """KL loss and d loss / d logits of the reference formula (train.py:190-255 as restated in
    oracle/fddm_torch_port.py) on the GPU in `dtype` (fp64: the checker; fp32: the reference's own arithmetic),
    over the whole batch, evaluated in batch chunks (the loss is a mean of per-sample terms)."""
    B = logits.shape[0]
    bt = betas.to(dtype)
    total = 0.0
    grads = []
    for i in range(0, B, chunk):
        sl = slice(i, min(B, i + chunk))
        lg = logits[sl].to(dtype).requires_grad_(True)
        part = P.kl_term(xt[sl], x0[sl], lg, t[sl], bt, mask[sl]) * ((sl.stop - sl.start) / B)
        part.backward()
        total += float(part.detach())
        grads.append(lg.grad.double())
    return total, torch.cat(grads, 0)


@pytest.mark.parametrize("tag,B,L,V,dtype", [("c2", 32, 128, 8000, torch.float32),
                                             ("c5shard", 64, 256, 8000, torch.float32),
                                             ("c4", 64, 256, 32000, torch.float32),
                                             ("c4", 64, 256, 32000, torch.bfloat16),
                                             ("c2", 32, 128, 8000, torch.bfloat16)])
def test_kl_fullsize_vs_port_fp64(fb, tag, B, L, V, dtype):
    """Whole-tensor check of the fused KL kernel (loss AND every gradient entry) against the fp64 port at
    the full BASELINE sizes: 1e-5 relative in fp32, 1e-2 with bf16 logits (north_star bars), in the max norm
    over the tensor.  Row by row (each row against ITS OWN largest entry) the kernel must additionally be as
    close to fp64 as the reference's own fp32 arithmetic is on that row: rows whose gradient is a cancellation
    residue (t <= 2, or a tiny posterior mismatch) are ill-conditioned in fp32 for the reference too."""
    s = sched(fb, V)
    ad = fb.SchedulerAdapter(s)
    logits, x0, mask, t = synth(B, L, V, dtype)
    xt = ad.sample_q(x0, t)
    lg = logits.clone().requires_grad_(True)
    loss = ad.kl_term(xt, x0, lg, t, mask)
    loss.backward()
    want, wgrad = port_kl(xt, x0, logits, t, s.betas, mask, torch.float64)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert abs(float(loss.detach()) - want) <= tol * abs(want), (float(loss.detach()), want)
    err = float((lg.grad.double() - wgrad).abs().max()) / float(wgrad.abs().max())
    assert err <= tol, err
    # row by row, each row against its own largest entry: rows with x_t == x_0 or t <= 2 have gradients that are
    # cancellation residues (1e-3..1e-5 of the tensor's largest entry) and are ill-conditioned in fp32 for the
    # reference as well (measured at c2: the reference's own fp32 arithmetic is up to 8e-3 away from fp64 on
    # such rows, this kernel up to 9e-3, medians 5e-7 vs 2e-7).  The kernel's per-row error DISTRIBUTION must
    # be no worse than twice the reference-fp32 one at every quantile.
    # (the reference's own arithmetic = the logits dtype: fp32 for fp32 logits; with bf16 logits the reference does
    #  its softmax in bf16, quirk Q11, and returns a bf16 gradient)
    _, ref32_grad = port_kl(xt, x0, logits, t, s.betas, mask, dtype)
    rmax = wgrad.abs().amax(-1).clamp_min(1e-300)
    ours = ((lg.grad.double() - wgrad).abs().amax(-1) / rmax)[mask]
    ref = ((ref32_grad - wgrad).abs().amax(-1) / rmax)[mask]
    q = torch.tensor([0.5, 0.9, 0.99, 0.999, 1.0], device="cuda", dtype=torch.float64)
    qo, qr = torch.quantile(ours, q), torch.quantile(ref, q)
    assert bool((qo <= torch.maximum(2.0 * qr, torch.full_like(qr, 2 * tol))).all()), (qo.tolist(), qr.tolist())


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_jumpy_sampler_c3_vs_port(fb, mode):
    """c3 (B=256 L=128 V=8000, T_infer=20 r=5), greedy: the ids every jump hands to the decoder equal the ids
    the torch port of the reference (fp32, same GPU) produces from the same inputs; a mismatch must be a
    near-tie of the port's posterior (top-2 within 1e-6 relative)."""
    B, L, V = 256, 128, 8000
    s = sched(fb, V)
    g = torch.Generator(device="cuda").manual_seed(7)
    steps = [(torch.randn(B, L, V, generator=g, device="cuda") * 3.0) for _ in range(4)]
    calls = []

    def decoder(x, t, c):
        calls.append((x.clone(), int(t[0])))
        return steps[len(calls) - 1]

    smp = fb.DiffusionJumpySampler(s, decoder, K=V, T_train=T_TRAIN, T_infer=20, r=5, greedy=True,
                                   sampling_mode=mode, device=torch.device("cuda"))
    x0, p_last = smp.sample(torch.zeros(B, 1, 1, device="cuda"), L)
    assert [c[1] for c in calls] == [20, 15, 10, 5]
    nbad = 0
    for i in range(4):
        x_t, t_s = calls[i]
        want, p_x0 = P.jump_once(x_t, steps[i], t_s, 5, s.betas, s.alpha_bar, V, T_TRAIN, 20, sampling_mode=mode, greedy=True)
        got = calls[i + 1][0] if i < 3 else smp.last_resampled_idx
        bad = want != got
        if bool(bad.any()):
            # re-evaluate the posterior rows in question to certify the near-tie
            if mode == "exact":
                oh = torch.zeros(B, L, V, device="cuda").scatter_(-1, x_t.unsqueeze(-1), 1.0)
                post = P.q_posterior_multi_step(oh, p_x0, torch.full((B,), t_s, device="cuda"), 5, s.betas, T_TRAIN)
            else:
                post = smp._alpha_bar_at_t_train(max(0, t_s - 5)) * p_x0 + (1 - smp._alpha_bar_at_t_train(max(0, t_s - 5))) / V
            top2 = post[bad].topk(2, -1).values
            assert float(((top2[:, 0] - top2[:, 1]) / top2[:, 0]).max()) < 1e-6
        nbad += int(bad.sum())
    assert nbad <= 2
    np.testing.assert_allclose(p_last.cpu().numpy(), torch.softmax(steps[3], -1).cpu().numpy(), rtol=1e-5, atol=1e-30)
    assert torch.equal(x0, p_last.argmax(-1))                               # sampler:292 (Q9)


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_jumpy_sampler_c3_fused_equals_dense(fb, mode):
    """c3: B=256 L=128 V=8000, T_infer=20 r=5.  Greedy ids of the fused jump kernel == argmax of the dense
    posterior built from this library's (oracle-checked) dense kernels + torch softmax; the whole sample()
    chain returns argmax of the last p_x0 and normalised probabilities."""
    B, L, V = 256, 128, 8000
    s = sched(fb, V)
    g = torch.Generator(device="cuda").manual_seed(7)
    steps = [(torch.randn(B, L, V, generator=g, device="cuda") * 3.0) for _ in range(4)]
    calls = []

    def decoder(x, t, c):
        calls.append((x.clone(), int(t[0])))
        return steps[len(calls) - 1]

    smp = fb.DiffusionJumpySampler(s, decoder, K=V, T_train=T_TRAIN, T_infer=20, r=5, greedy=True,
                                   sampling_mode=mode, device=torch.device("cuda"))
    x0, p_last = smp.sample(torch.zeros(B, 1, 1, device="cuda"), L)
    assert [c[1] for c in calls] == [20, 15, 10, 5]
    assert x0.shape == (B, L) and p_last.shape == (B, L, V)
    assert torch.equal(x0, p_last.argmax(-1))                               # sampler:292
    assert float((p_last.sum(-1) - 1).abs().max()) < 1e-5
    assert float((p_last - torch.softmax(steps[3], -1)).abs().max()) < 1e-6
    # every transition: ids seen by the decoder at jump i+1 == argmax of the dense posterior of jump i
    for i in range(3):
        x_t, t_s = calls[i]
        p = torch.softmax(steps[i], -1)
        if mode == "exact":
            oh = torch.zeros(B, L, V, device="cuda").scatter_(-1, x_t.unsqueeze(-1), 1.0)
            post = s.q_posterior_multi_step(oh, p, torch.full((B,), t_s, device="cuda"), 5)
        else:
            ab = smp._alpha_bar_at_t_train(max(0, t_s - 5))
            post = ab * p + (1.0 - ab) * (1.0 / V)
        want = post.argmax(-1)
        got = calls[i + 1][0]
        bad = (want != got)
        if bool(bad.any()):                                                  # only exact near-ties may differ
            top2 = post[bad].topk(2, -1).values
            assert float(((top2[:, 0] - top2[:, 1]) / top2[:, 0]).max()) < 1e-6
        assert int(bad.sum()) <= 2


def test_jump_sampling_c3_statistics(fb):
    """c3 with Categorical resampling (in-kernel RNG): ids in range, and the empirical probability of
    resampling the argmax token matches the mean posterior mass of that token."""
    B, L, V = 256, 128, 8000
    s = sched(fb, V)
    g = torch.Generator(device="cuda").manual_seed(3)
    logits = torch.randn(B, L, V, generator=g, device="cuda") * 4.0
    x_t = torch.randint(0, V, (B, L), generator=g, device="cuda")
    smp = fb.DiffusionJumpySampler(s, lambda x, t, c: logits, K=V, T_train=T_TRAIN, T_infer=20, r=5, greedy=False,
                                   sampling_mode="exact", device=torch.device("cuda"))
    smp.generator = torch.Generator(device="cuda"); smp.generator.manual_seed(5)
    ids, _ = smp._jump_once(x_t, 5, 5, torch.zeros(B, 1, 1, device="cuda"), L)
    assert int(ids.min()) >= 0 and int(ids.max()) < V
    oh = torch.zeros(B, L, V, device="cuda").scatter_(-1, x_t.unsqueeze(-1), 1.0)
    post = s.q_posterior_multi_step(oh, torch.softmax(logits, -1), torch.full((B,), 5, device="cuda"), 5)
    top = post.argmax(-1)
    p_top = post.gather(-1, top.unsqueeze(-1)).squeeze(-1).double()
    n = B * L
    hit = float((ids == top).double().mean())
    mean, sd = float(p_top.mean()), float((p_top * (1 - p_top)).sum().sqrt() / n)
    assert abs(hit - mean) < 5 * sd + 1e-4, (hit, mean, sd)


def test_sample_q_fullsize_keep_rate(fb):
    B, L, V = 64, 256, 8000
    s = sched(fb, V)
    _, x0, _, t = synth(B, L, V, torch.float32)
    xt = s.sample_q_ids(x0, t)
    ab = s.alpha_bar[t - 1].double()
    p_keep = (ab + (1 - ab) / V)[:, None].expand(B, L)
    hit = (xt == x0).double()
    z = (hit.sum() - p_keep.sum()) / (p_keep * (1 - p_keep)).sum().sqrt()
    assert abs(float(z)) < 5.0
    moved = xt[xt != x0]
    assert int(moved.min()) >= 0 and int(moved.max()) < V


@pytest.mark.parametrize("B,T,D,dtype", [(32, 128, 768, torch.float32), (64, 256, 768, torch.float32),
                                         (32, 128, 256, torch.float32), (64, 256, 768, torch.bfloat16)])
def test_lfd_fullsize_properties(fb, B, T, D, dtype):
    g = torch.Generator(device="cuda").manual_seed(11)
    za = torch.randn(B, T, D, generator=g, device="cuda")
    zb = 0.9 * za + 0.436 * torch.randn(B, T, D, generator=g, device="cuda")
    za, zb = za.to(dtype), zb.to(dtype)
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    with torch.no_grad():
        l_ab = float(fb.lfd_loss(za, zb, 5e-3))
        l_ba = float(fb.lfd_loss(zb, za, 5e-3))                              # C -> C^T: same loss
        l_aff = float(fb.lfd_loss(za * 3.0 + 1.5, zb * 4.0 - 2.0, 5e-3))     # standardisation: affine invariance
        l_self = float(fb.lfd_loss(za, za, 0.0))                              # C_jj = var/(var+eps) ~ 1
    assert abs(l_ab - l_ba) <= tol * abs(l_ab)
    # (only approximate: eps=1e-5 sits inside sqrt(var+eps), so rescaling shifts C_jj by ~eps/var)
    assert abs(l_ab - l_aff) <= max(tol, 5e-4) * abs(l_ab)
    assert 0 <= l_self < D * 1e-6
    # torch fp64 evaluation of the same formula (losses:18-58) on the GPU
    def ref(a, b):
        a = a.double(); b = b.double()
        sa = (a - a.mean(0, keepdim=True)) / torch.sqrt(a.var(0, unbiased=False, keepdim=True) + 1e-5)
        sb = (b - b.mean(0, keepdim=True)) / torch.sqrt(b.var(0, unbiased=False, keepdim=True) + 1e-5)
        C = sa.reshape(-1, D).T @ sb.reshape(-1, D) / (B * T)
        d = torch.diagonal(C)
        return ((1 - d) ** 2).sum() + 5e-3 * ((C - torch.diag(d)) ** 2).sum()
    a64 = za.detach().clone().requires_grad_(True); b64 = zb.detach().clone().requires_grad_(True)
    want = ref(a64, b64)
    want.backward()
    a = za.detach().clone().requires_grad_(True); b = zb.detach().clone().requires_grad_(True)
    got = fb.lfd_loss(a, b, 5e-3)
    got.backward()
    assert abs(float(got) - float(want)) <= tol * abs(float(want))
    gtol = 2e-5 if dtype == torch.float32 else 2e-2
    for x, y in ((a.grad, a64.grad), (b.grad, b64.grad)):
        assert float((x.double() - y.double()).abs().max()) <= gtol * float(y.double().abs().max())
