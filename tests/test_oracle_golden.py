"""Pins the numpy oracle (oracle/fddm_oracle.py) against vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import fddm_oracle as O

RTOL = 2e-6      # fp32 restatement vs torch fp32 (different exp/sum implementations)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_tables(golden):
    b, ab = O.make_tables(8000, 200, 0.2)
    assert rel_err(b, golden["tab_betas"]) < 1e-6
    assert rel_err(ab, golden["tab_alpha_bar"]) < 1e-5
    b, ab = O.make_tables(53, 50, 0.3)
    assert rel_err(b, golden["tab53_betas"]) < 1e-6
    assert rel_err(ab, golden["tab53_alpha_bar"]) < 1e-5


def test_q_sample_onehot_clamp(golden):
    ids = golden["qs_x0_ids"]; t = golden["qs_t"]
    oh = np.zeros(ids.shape + (8000,), np.float32)
    np.put_along_axis(oh, ids[..., None], 1.0, -1)
    got = O.q_sample(oh, t, golden["tab_alpha_bar"])
    ref = golden["qs_out"]
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=0)
    np.testing.assert_allclose(got.sum(-1), 1.0, atol=1e-5)       # the reference's own sanity check
    # quirk Q4: at t=1 the uniform floor (1-abar)/K is below eps, so the clamp changed the value
    lo = got[0, 0, (ids[0, 0] + 1) % 8000]
    assert lo > (1.0 - golden["tab_alpha_bar"][0]) / 8000 * 1.5


def test_q_sample_soft(golden):
    got = O.q_sample(golden["qs53_x0"], golden["qs53_t"], golden["tab53_alpha_bar"])
    np.testing.assert_allclose(got, golden["qs53_out"], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("tag,K,tab", [("sq", 8000, "tab_alpha_bar"), ("sq53", 53, "tab53_alpha_bar")])
def test_sample_q_ids_bit_exact(golden, tag, K, tab):
    x0 = golden[f"{tag}_x0"]; B, L = x0.shape
    got = O.sample_q_ids(x0, golden[f"{tag}_t"], golden[tab], K, golden[f"{tag}_E"].reshape(B, L, K))
    assert np.array_equal(got, golden[f"{tag}_xt"])


def test_q_posterior(golden):
    got = O.q_posterior(golden["qp_xt"], golden["qp_xh"], golden["qp_t"], golden["tab53_betas"])
    np.testing.assert_allclose(got, golden["qp_out"], rtol=5e-6, atol=1e-12)
    np.testing.assert_allclose(got.sum(-1), 1.0, atol=1e-5)


def test_q_posterior_multi_step(golden):
    for ci in range(int(golden["ms_n"])):
        got = O.q_posterior_multi_step(golden["qp_xt"], golden["qp_xh"], golden[f"ms{ci}_t"],
                                       int(golden[f"ms{ci}_delta"]), golden["tab53_betas"], 50)
        np.testing.assert_allclose(got, golden[f"ms{ci}_out"], rtol=5e-6, atol=1e-12, err_msg=f"case {ci}")


def test_q_posterior_multi_step_8k_aliased_recurrence(golden):
    xt = np.zeros((3, 1, 8000), np.float32); xt[:, 0, int(golden["ms8k_xt_id"])] = 1
    for ci in range(3):
        got = O.q_posterior_multi_step(xt, golden["ms8k_xh"], golden[f"ms8k{ci}_t"],
                                       int(golden[f"ms8k{ci}_delta"]), golden["tab_betas"], 200)
        np.testing.assert_allclose(got, golden[f"ms8k{ci}_out"], rtol=5e-6, atol=1e-14, err_msg=f"case {ci}")


def test_multistep_coeffs_differ_from_textbook(golden):
    # quirk Q1: the executed recurrence uses the updated a'; the textbook one would not.
    betas = golden["tab_betas"]
    _, a, b, _, _ = O.multistep_coeffs(np.array([200]), 5, betas, 8000, 200)
    a_t, b_t = np.float32(1), np.float32(0)
    for s in range(200, 195, -1):
        a_s = np.float32(1) - betas[s - 1]; b_s = betas[s - 1] / np.float32(8000)
        a_t, b_t = a_s * a_t, a_s * b_t + b_s * (a_t + np.float32(8000) * b_t)
    assert abs(a[0] - a_t) < 1e-7
    assert abs(b[0] - b_t) / b_t > 0.05


@pytest.mark.parametrize("tag,io,tab,tol", [
    ("f32", "f32", "tab53_betas", 3e-6), ("f32m", "f32", "tab53_betas", 3e-6),
    ("bf16m", "bf16", "tab53_betas", 1e-2), ("f16", "f16", "tab53_betas", 2e-3),
    ("f32k8", "f32", "tab_betas", 3e-6)])
def test_kl_term_and_grad(golden, tag, io, tab, tol):
    mask = golden[f"kl_{tag}_mask"]
    mask = None if mask.size == 0 else mask
    loss, grad = O.kl_term(golden[f"kl_{tag}_xt"], golden[f"kl_{tag}_x0"], golden[f"kl_{tag}_logits"],
                           golden[f"kl_{tag}_t"], golden[tab], mask, io_dtype=io, want_grad=True)
    assert str(golden[f"kl_{tag}_loss_dtype"]) == "torch.float32"          # always fp32 (survey a7)
    assert abs(float(loss) - float(golden[f"kl_{tag}_loss"])) <= tol * abs(float(golden[f"kl_{tag}_loss"]))
    assert rel_err(grad, golden[f"kl_{tag}_grad"]) < max(tol, 1e-5)
    # fp64 tie-breaker agrees too
    loss64, grad64 = O.kl_term(golden[f"kl_{tag}_xt"], golden[f"kl_{tag}_x0"], golden[f"kl_{tag}_logits"],
                               golden[f"kl_{tag}_t"], golden[tab], mask, io_dtype=io, dtype=np.float64,
                               want_grad=True)
    assert abs(float(loss64) - float(golden[f"kl_{tag}_loss"])) <= max(tol, 1e-5) * abs(float(loss64))
    if io == "f32":
        assert rel_err(grad64, golden[f"kl_{tag}_grad"]) < 1e-5


def test_kl_empty_sample_contributes_zero(golden):
    # quirk Q6: sample 1 of the masked fixtures has no valid token
    mask = golden["kl_f32m_mask"]
    assert mask[1].sum() == 0
    tok, _ = O.kl_token_terms(golden["kl_f32m_xt"], golden["kl_f32m_x0"], golden["kl_f32m_logits"],
                              golden["kl_f32m_t"], golden["tab53_betas"])
    loss, w = O.kl_reduce(tok, mask)
    assert np.all(w[1] == 0) and np.isfinite(loss)


@pytest.mark.parametrize("tag,tol", [("r0", 2e-6), ("r9", 2e-5), ("l1", 2e-6)])
def test_lfd_loss_and_grads(golden, tag, tol):
    loss, ga, gb = O.lfd_loss(golden[f"lfd_{tag}_za"], golden[f"lfd_{tag}_zb"],
                              float(golden[f"lfd_{tag}_lam"]), want_grad=True)
    ref = float(golden[f"lfd_{tag}_loss"])
    assert abs(float(loss) - ref) <= tol * abs(ref)
    assert rel_err(ga, golden[f"lfd_{tag}_ga"]) < 2e-5
    assert rel_err(gb, golden[f"lfd_{tag}_gb"]) < 2e-5
    loss64, ga64, gb64 = O.lfd_loss(golden[f"lfd_{tag}_za"], golden[f"lfd_{tag}_zb"],
                                    float(golden[f"lfd_{tag}_lam"]), dtype=np.float64, want_grad=True)
    assert abs(float(loss64) - ref) <= 1e-5 * abs(ref)
    assert rel_err(ga64, golden[f"lfd_{tag}_ga"]) < 1e-5


def test_lfd_bf16_within_1e2(golden):
    assert str(golden["lfd_bf16_loss_dtype"]) == "torch.bfloat16"          # output in input dtype
    loss, _, _ = O.lfd_loss(golden["lfd_bf16_za"], golden["lfd_bf16_zb"], float(golden["lfd_bf16_lam"]))
    ref = float(golden["lfd_bf16_loss"])
    assert abs(float(loss) - ref) <= 2e-2 * abs(ref)


SAMPLER_TAGS = ["ex_g", "ex_s", "ex_map", "ex_max", "fa_g", "fa_s", "fa_st", "ex_st", "ex_s_bf16",
                "fa_g_bf16", "ex_s_8k"]


@pytest.mark.parametrize("tag", SAMPLER_TAGS)
def test_jumpy_sampler_chain(golden, tag):
    K, T_train, T_infer, r, B, L, greedy = [int(v) for v in golden[f"js_{tag}_cfg"]]
    mode, pmode, temp, dt = [str(s) for s in golden[f"js_{tag}_mode"]]
    io = {"torch.float32": "f32", "torch.bfloat16": "bf16", "torch.float16": "f16"}[dt]
    logits = golden[f"js_{tag}_logits"]; noise = golden[f"js_{tag}_noise"]
    betas = golden["tab_betas"] if K == 8000 else golden["tab53_betas"]
    abar = golden["tab_alpha_bar"] if K == 8000 else golden["tab53_alpha_bar"]
    seen = []

    def logits_fn(x, t, step):
        seen.append(x.copy())
        return logits[step]

    noise_fn = (lambda s: noise[s].reshape(B, L, K)) if noise.size else None
    x0, p_last, _ = O.sample_loop(golden[f"js_{tag}_xT"], logits_fn, K=K, T_train=T_train, T_infer=T_infer,
                                  r=r, betas=betas, alpha_bar=abar, sampling_mode=mode, posterior_mode=pmode,
                                  greedy=bool(greedy), temperature=float(temp), noise_fn=noise_fn, io_dtype=io)
    # the ids fed to the decoder at every jump are the reference's, bit for bit
    assert np.array_equal(np.stack(seen), golden[f"js_{tag}_x_seen"])
    assert np.array_equal(x0, golden[f"js_{tag}_x0"])
    tol = 1e-2 if io != "f32" else 2e-6
    np.testing.assert_allclose(p_last, golden[f"js_{tag}_p_last"], rtol=tol, atol=1e-30)
    assert np.all(golden[f"js_{tag}_t_seen"][:, 0] == np.arange(T_infer, 0, -r)[: len(seen)])


def test_fast_mode_alpha_bar_off_by_one(golden):
    # quirk Q3: 1-based train index on a 0-based table
    abar = golden["tab_alpha_bar"]
    assert O.alpha_bar_at_t_train(10, 20, 200, abar) == abar[100]
    assert O.alpha_bar_at_t_train(0, 20, 200, abar) == 1.0
    with pytest.raises(IndexError):
        O.alpha_bar_at_t_train(20, 20, 200, abar)


def test_cer_wer_oracle_matches_reference_golden():
    """the oracle's Levenshtein CER / WER equal the reference's (models/evaluate.py:94-134) on the
    committed pairs generated by tests/golden/make_cer_golden.py (CJK + Latin, empty and identical strings)."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cer_wer_vectors.json")
    vec = json.load(open(path, encoding="utf-8"))
    assert len(vec) >= 60
    for v in vec:
        assert O.calculate_cer(v["ref"], v["hyp"]) == v["cer"]
        assert O.calculate_wer(v["ref"], v["hyp"]) == v["wer"]
