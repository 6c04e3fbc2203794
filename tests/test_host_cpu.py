"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header declares, the
host mirror's argument checking / control logic, the torch-eager CPU port against the golden vectors,
and the batch-sharded decomposition (world_size 2, gloo)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import fddm_oracle as O
from oracle import fddm_torch_port as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ------------------------------------------------------------------------------------------------
# the C-ABI library
# ------------------------------------------------------------------------------------------------
def header_symbols():
    text = open(os.path.join(ROOT, "include", "fddm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fddm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import fddm_b200
    lib = ctypes.CDLL(fddm_b200._lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/fddm_b200.h but not exported"
    assert fddm_b200._lib.MISSING == []
    assert set(fddm_b200._lib.SIGNATURES) == set(names)          # the ctypes table binds exactly the header


def header_prototypes():
    """name -> number of parameters, parsed from the C prototypes in include/fddm_b200.h"""
    text = open(os.path.join(ROOT, "include", "fddm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(fddm_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_ctypes_signatures_match_header_arity():
    """A ctypes argtypes list that disagrees with the C prototype corrupts the call silently: the binding
    table must have exactly as many parameters as each prototype, pointers where the header has pointers."""
    import fddm_b200
    protos = header_prototypes()
    assert set(protos) == set(fddm_b200._lib.SIGNATURES)
    for name, (_, argtypes) in fddm_b200._lib.SIGNATURES.items():
        assert len(argtypes) == protos[name], f"{name}: header has {protos[name]} parameters, ctypes table {len(argtypes)}"
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "fddm_b200.h")).read(), flags=re.S)
    for m in re.finditer(r"\b(fddm_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), [p.strip() for p in m.group(2).split(",")]
        if params in ([""], ["void"]):
            continue
        argtypes = fddm_b200._lib.SIGNATURES[name][1]
        for i, (prm, at) in enumerate(zip(params, argtypes)):
            is_ptr = "*" in prm or "fddm_stream_t" in prm
            assert is_ptr == (at is ctypes.c_void_p), f"{name} parameter {i} ({prm!r}) vs ctypes {at}"
            if not is_ptr:
                want = {"int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "int": ctypes.c_int, "float": ctypes.c_float,
                        "double": ctypes.c_double}[prm.split()[0]]
                assert at is want, f"{name} parameter {i} ({prm!r}) vs ctypes {at}"


def test_dropin_shims_resolve_to_the_b200_implementation():
    """fddm-asr_b200/dropin mirrors the reference's import paths."""
    import subprocess
    import sys
    code = ("from fddm.sched.diffusion_scheduler import DiscreteDiffusionScheduler as S;"
            "from losses.fddm_losses import lfd_loss;"
            "from sampler.jumpy_sampler import DiffusionJumpySampler, ModelAdapter;"
            "print(S.__module__, lfd_loss.__module__, DiffusionJumpySampler.__module__, ModelAdapter.__module__)")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "fddm-asr_b200", "dropin"))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["fddm_b200.scheduler", "fddm_b200.losses", "fddm_b200.sampler", "fddm_b200.sampler"]


def test_host_only_entry_points():
    import fddm_b200
    lib = fddm_b200._lib.lib
    assert lib.fddm_version() == 3
    assert lib.fddm_kl_workspace_bytes(32, 128) >= 128 + 32 * 128 * 4
    assert lib.fddm_kl_workspace_bytes(0, 5) == 0
    w = lib.fddm_lfd_workspace_bytes(32, 128, 768)
    assert w >= 2 * 32 * 128 * 768 * 4 + 4 * 128 * 768 * 4
    assert lib.fddm_lfd_workspace_bytes(32, 0, 768) == 0
    # argument validation happens before any CUDA call: null pointers -> EINVAL + message
    rc = lib.fddm_q_sample_dense(None, None, None, 10, 1, 1, 8, 1e-8, None, None)
    assert rc == -1 and b"null" in lib.fddm_last_error()
    # L_fd row order: B < 32 keeps the natural order (one partial per (t, d)); B >= 32 is tb-major with the batch
    # padded to a multiple of 32 and one partial per 32 batch rows
    assert [lib.fddm_lfd_bn_parts(B, 7, 64) for B in (1, 4, 31, 32, 33, 64, 130, 512)] == [1, 1, 1, 1, 2, 2, 5, 16]
    assert lib.fddm_lfd_workspace_bytes(33, 128, 768) > lib.fddm_lfd_workspace_bytes(32, 128, 768)   # padded planes
    assert lib.fddm_jump_workspace_bytes(8, 64) == 128 + 8 * 64 * 4 and lib.fddm_jump_workspace_bytes(0, 64) == 0
    assert lib.fddm_edit_distance_workspace_bytes(10, 99) == 10 * 100 * 4
    assert lib.fddm_set_sm_reserve(16) == 0 and lib.fddm_set_sm_reserve(0) == 0
    assert lib.fddm_set_sm_reserve(-1) == -1 and b"set_sm_reserve" in lib.fddm_last_error()
    rc = lib.fddm_edit_distance(None, None, None, None, 1, 4, None, None, None)
    assert rc == -1 and b"null" in lib.fddm_last_error()
    # the library's own all-reduce: pad size, and argument validation before any CUDA call
    assert lib.fddm_xgpu_signal_pad_bytes() == 4096 + (32 * 16 + 32) * 4
    L = fddm_b200._lib
    assert lib.fddm_xgpu_allreduce(None, None, None, 0, 2, 4, 16, L.XGPU_P2P, 0, None) == -1          # no signal pads
    assert lib.fddm_xgpu_allreduce(8, None, 8, 0, 1, 4, 16, L.XGPU_P2P, 0, None) == -1 and b"world" in lib.fddm_last_error()
    assert lib.fddm_xgpu_allreduce(8, None, 8, 0, 2, 4, 18, L.XGPU_P2P, 0, None) == -1 and b"16 bytes" in lib.fddm_last_error()
    assert lib.fddm_xgpu_allreduce(8, None, 8, 0, 2, 2, 16, L.XGPU_P2P, 0, None) == -1 and b"elem_bytes" in lib.fddm_last_error()
    assert lib.fddm_xgpu_allreduce(None, None, 8, 0, 2, 8, 16, L.XGPU_NVLS, 0, None) == -1 and b"multicast" in lib.fddm_last_error()
    assert lib.fddm_xgpu_allreduce(8, None, 8, 0, 2, 8, 16, 7, 0, None) == -1 and b"algorithm" in lib.fddm_last_error()
    assert not fddm_b200.symmetric_exchange_available()              # no process group here
    # the measurement aid works without a device: nothing recorded -> empty report
    fddm_b200._lib.profile_enable(True)
    assert fddm_b200._lib.profile_read() == {}
    fddm_b200._lib.profile_enable(False)


def test_no_undefined_names_in_host_code():
    """The multi-GPU branches of the host mirror and of bench.py only execute on a multi-GPU box; a name that a
    refactor dropped would first be noticed there.  A small AST walk (names loaded but never bound anywhere in the
    module, builtins aside) catches that class of mistake here."""
    import ast
    import builtins
    import glob
    files = (glob.glob(os.path.join(ROOT, "fddm-asr_b200", "fddm_b200", "*.py")) + glob.glob(os.path.join(ROOT, "scripts", "*.py"))
             + glob.glob(os.path.join(ROOT, "fddm-asr_b200", "dropin", "**", "*.py"), recursive=True)
             + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")])
    assert len(files) > 12
    for path in files:
        tree = ast.parse(open(path).read())
        bound = set(dir(builtins)) | {"__file__"}
        for node in ast.walk(tree):
            if isinstance(node, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
                bound.add(node.name)
            elif isinstance(node, ast.Import):
                bound.update((a.asname or a.name).split(".")[0] for a in node.names)
            elif isinstance(node, ast.ImportFrom):
                bound.update(a.asname or a.name for a in node.names)
            elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
                bound.add(node.id)
            elif isinstance(node, ast.arg):
                bound.add(node.arg)
            elif isinstance(node, ast.ExceptHandler) and node.name:
                bound.add(node.name)
        missing = sorted({(n.id, n.lineno) for n in ast.walk(tree)
                          if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in bound})
        assert not missing, f"{path}: names used but never bound: {missing}"


def test_metrics_host_logic_rejects_cpu_only_runs():
    """calculate_cer / calculate_wer mirror models/evaluate.py:94-134; like every op of the package they have no
    CPU fallback (the oracle restatement lives in oracle/fddm_oracle.py and is pinned by test_oracle_golden.py)."""
    import fddm_b200 as fb
    assert fb.batch_cer([], []) == [] and fb.batch_wer([], []) == []
    with pytest.raises(ValueError):
        fb.batch_cer(["a"], ["a", "b"])
    if not torch.cuda.is_available():
        with pytest.raises(ValueError):
            fb.calculate_cer("abc", "abd")


# ------------------------------------------------------------------------------------------------
# host mirror: construction, validation, control flow (no compute)
# ------------------------------------------------------------------------------------------------
def test_scheduler_tables_and_cpu_rejection(golden):
    import fddm_b200 as fb
    s = fb.DiscreteDiffusionScheduler(K=8000, T=200, device=torch.device("cpu"))
    assert (s.K, s.T, s.eps) == (8000, 200, 1e-8)
    assert np.array_equal(s.betas.numpy(), golden["tab_betas"])          # same torch ops, same device => bit-equal
    assert np.array_equal(s.alpha_bar.numpy(), golden["tab_alpha_bar"])
    assert s.w_prefix is s.alpha_bar
    with pytest.raises(ValueError, match="CUDA tensors only"):
        s.q_sample(torch.zeros(1, 2, 8000), torch.ones(1, dtype=torch.long))
    with pytest.raises(AssertionError):
        s.q_posterior(torch.zeros(1, 2, 7), torch.zeros(1, 2, 7), torch.ones(1, dtype=torch.long))
    ad = fb.SchedulerAdapter(s)
    assert ad.sch is s
    t = torch.tensor([1, 200, 17])
    assert torch.equal(ad.w_t(t), s.alpha_bar[t - 1])
    with pytest.raises(ValueError):
        ad.kl_term(torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 2, 8000),
                   torch.ones(1, dtype=torch.long))
    with pytest.raises(ValueError):
        fb.lfd_loss(torch.zeros(2, 3, 8), torch.zeros(2, 3, 8))
    with pytest.raises(AssertionError):
        fb.lfd_loss(torch.zeros(2, 3, 8), torch.zeros(2, 3, 9))


def test_sampler_host_logic(golden):
    import fddm_b200 as fb
    s = fb.DiscreteDiffusionScheduler(K=8000, T=200, device=torch.device("cpu"))
    smp = fb.DiffusionJumpySampler(s, decoder=None, K=8000, T_train=200, T_infer=20, r=5, greedy=False,
                                   posterior_mode="map", sampling_mode="fast", temperature=0.7,
                                   device=torch.device("cpu"))
    assert smp._alpha_bar_index(10) == 100                               # quirk Q3: 1-based index on a 0-based table
    assert float(smp._alpha_bar_at_t_train(10)) == float(golden["tab_alpha_bar"][100])
    assert float(smp._alpha_bar_at_t_train(0)) == 1.0
    with pytest.raises(IndexError):
        smp._alpha_bar_at_t_train(20)
    for ti in range(0, 20):
        want = O.alpha_bar_at_t_train(ti, 20, 200, golden["tab_alpha_bar"])
        assert float(smp._alpha_bar_at_t_train(ti)) == float(want)
    assert smp.get_sampling_info() == {"sampling_mode": "fast", "posterior_mode": "map", "T_infer": 20, "r": 5,
                                       "greedy": False, "temperature": 0.7, "K": 8000}

    class NoTables:
        pass
    with pytest.raises(ValueError):
        fb.DiffusionJumpySampler(NoTables(), None, K=10, T_train=5, T_infer=5)


# ------------------------------------------------------------------------------------------------
# the torch-eager CPU port (bench.py's CPU arm) against the reference-generated vectors
# ------------------------------------------------------------------------------------------------
def test_port_tables(golden):
    b, ab = P.make_tables(8000, 200)
    assert np.array_equal(b.numpy(), golden["tab_betas"]) and np.array_equal(ab.numpy(), golden["tab_alpha_bar"])


def _inject(monkeypatch, E):
    """torch.multinomial(p, 1) == argmax(p / E) with E ~ Exp(1): replay the recorded E."""
    it = iter(E)
    monkeypatch.setattr(torch, "multinomial", lambda p, n, replacement=False, **kw: (p / next(it)).argmax(-1, keepdim=True))


@pytest.mark.parametrize("tag,K,T,bm", [("sq", 8000, 200, 0.2), ("sq53", 53, 50, 0.3)])
def test_port_sample_q(golden, monkeypatch, tag, K, T, bm):
    _, ab = P.make_tables(K, T, bm)
    _inject(monkeypatch, [torch.from_numpy(golden[f"{tag}_E"])])
    got = P.sample_q(torch.from_numpy(golden[f"{tag}_x0"]), torch.from_numpy(golden[f"{tag}_t"]), ab, K)
    assert np.array_equal(got.numpy(), golden[f"{tag}_xt"])


@pytest.mark.parametrize("tag,K,T,bm", [("f32", 53, 50, 0.3), ("f32m", 53, 50, 0.3), ("f32k8", 8000, 200, 0.2)])
def test_port_kl(golden, tag, K, T, bm):
    betas, _ = P.make_tables(K, T, bm)
    mask = golden[f"kl_{tag}_mask"]
    mask = None if mask.size == 0 else torch.from_numpy(mask)
    lg = torch.from_numpy(golden[f"kl_{tag}_logits"]).requires_grad_(True)
    loss = P.kl_term(torch.from_numpy(golden[f"kl_{tag}_xt"]), torch.from_numpy(golden[f"kl_{tag}_x0"]), lg,
                     torch.from_numpy(golden[f"kl_{tag}_t"]), betas, mask)
    loss.backward()
    assert abs(float(loss.detach()) - float(golden[f"kl_{tag}_loss"])) <= 1e-6 * abs(float(golden[f"kl_{tag}_loss"]))
    assert rel_err(lg.grad.numpy(), golden[f"kl_{tag}_grad"]) < 1e-6


@pytest.mark.parametrize("tag", ["r0", "r9", "l1"])
def test_port_lfd(golden, tag):
    a = torch.from_numpy(golden[f"lfd_{tag}_za"]).requires_grad_(True)
    b = torch.from_numpy(golden[f"lfd_{tag}_zb"]).requires_grad_(True)
    loss = P.lfd_loss(a, b, float(golden[f"lfd_{tag}_lam"]))
    loss.backward()
    assert abs(float(loss.detach()) - float(golden[f"lfd_{tag}_loss"])) <= 1e-6 * abs(float(golden[f"lfd_{tag}_loss"]))
    assert rel_err(a.grad.numpy(), golden[f"lfd_{tag}_ga"]) < 1e-5


@pytest.mark.parametrize("tag", ["ex_g", "ex_s", "fa_g", "fa_s", "ex_st"])
def test_port_jump_chain(golden, monkeypatch, tag):
    K, T_train, T_infer, r, B, L, greedy = [int(v) for v in golden[f"js_{tag}_cfg"]]
    mode, pmode, temp, dt = [str(s) for s in golden[f"js_{tag}_mode"]]
    betas, ab = P.make_tables(K, T_train, 0.2 if K == 8000 else 0.3)
    noise = golden[f"js_{tag}_noise"]
    if noise.size:
        _inject(monkeypatch, [torch.from_numpy(n) for n in noise])
    x = torch.from_numpy(golden[f"js_{tag}_xT"])
    t, step = T_infer, 0
    while t > 0:
        delta = min(r, t)
        assert np.array_equal(x.numpy(), golden[f"js_{tag}_x_seen"][step])
        x, p = P.jump_once(x, torch.from_numpy(golden[f"js_{tag}_logits"][step]), t, delta, betas, ab, K, T_train,
                           T_infer, sampling_mode=mode, greedy=bool(greedy), temperature=float(temp))
        t -= delta; step += 1
    assert np.array_equal(p.argmax(-1).numpy(), golden[f"js_{tag}_x0"])


# ------------------------------------------------------------------------------------------------
# batch sharding (world_size 2, gloo): the phase decomposition the C-ABI exposes -- per-rank partial
# results + SUM all-reduces -- reproduces the global-batch reference values.
# ------------------------------------------------------------------------------------------------
def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(99)
        B, L, V, T, D = 6, 5, 53, 50, 16
        betas, _ = O.make_tables(V, T, 0.3)
        x0 = rng.integers(0, V, size=(B, L)); xt = rng.integers(0, V, size=(B, L))
        t = rng.integers(1, T + 1, size=B); logits = rng.normal(size=(B, L, V)).astype(np.float32)
        mask = rng.random((B, L)) < 0.7
        za = rng.normal(size=(B, L, D)); zb = 0.8 * za + 0.6 * rng.normal(size=(B, L, D))
        lo, hi = rank * B // world, (rank + 1) * B // world
        sl = slice(lo, hi)
        red = lambda a: (lambda tt: (dist.all_reduce(tt), tt.numpy())[1])(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)))

        # KL: each rank reduces its own samples with the GLOBAL batch divisor; one scalar all-reduce
        tok, g = O.kl_token_terms(xt[sl], x0[sl], logits[sl], t[sl], betas, dtype=np.float64, want_grad=True)
        part, wgt = O.kl_reduce(tok, mask[sl], dtype=np.float64)
        nloc = hi - lo
        kl_global = red(np.array([part * nloc / B]))[0]
        grad_local = g * (wgt * nloc / B)[..., None]
        want_loss, want_grad = O.kl_term(xt, x0, logits, t, betas, mask, dtype=np.float64, want_grad=True)
        assert abs(kl_global - want_loss) < 1e-12 * abs(want_loss)
        assert np.allclose(grad_local, want_grad[sl], rtol=1e-12, atol=1e-18)

        # L_fd: stats -> all-reduce -> partial covariance -> all-reduce -> loss/G -> bn sums -> all-reduce
        a, b = za[sl], zb[sl]
        sums = red(np.stack([a.sum(0), (a * a).sum(0), b.sum(0), (b * b).sum(0)]))
        mean_a, mean_b = sums[0] / B, sums[2] / B
        rstd_a = 1 / np.sqrt(sums[1] / B - mean_a ** 2 + 1e-5); rstd_b = 1 / np.sqrt(sums[3] / B - mean_b ** 2 + 1e-5)
        at, bt = (a - mean_a) * rstd_a, (b - mean_b) * rstd_b
        N = B * L
        cov = red(at.reshape(-1, D).T @ bt.reshape(-1, D))
        C = cov / N
        dg = np.diagonal(C)
        loss = ((1 - dg) ** 2).sum() + 5e-3 * ((C - np.diag(dg)) ** 2).sum()
        G = 2 * 5e-3 * (C - np.diag(dg)); G[np.arange(D), np.arange(D)] = -2 * (1 - dg)
        dza = (bt.reshape(-1, D) @ G.T / N).reshape(a.shape); dzb = (at.reshape(-1, D) @ G / N).reshape(a.shape)
        local_m1 = np.stack([dza.sum(0), dzb.sum(0)])
        bn = red(np.stack([dza.sum(0), (dza * at).sum(0), dzb.sum(0), (dzb * bt).sum(0)]))
        da = (dza - bn[0] / B - at * bn[1] / B) * rstd_a
        db = (dzb - bn[2] / B - bt * bn[3] / B) * rstd_b
        wl, wa, wb = O.lfd_loss(za, zb, 5e-3, dtype=np.float64, want_grad=True)
        assert abs(loss - wl) < 1e-10 * abs(wl)
        assert np.allclose(da, wa[sl], rtol=1e-8, atol=1e-14) and np.allclose(db, wb[sl], rtol=1e-8, atol=1e-14)
        # what fddm_b200.losses does when sharded: only sum_b dz~*z~ is all-reduced; sum_b dz~ is zero over the
        # GLOBAL batch (z~ has zero global batch mean) -- though not per rank -- and is replaced by 0
        assert np.abs(bn[0]).max() < 1e-12 * np.abs(dza).max() and np.abs(bn[2]).max() < 1e-12 * np.abs(dzb).max()
        assert np.abs(local_m1).max() > 1e-3 * np.abs(dza).max()          # a rank's partial sum is far from zero
        da0 = (dza - at * bn[1] / B) * rstd_a
        db0 = (dzb - bt * bn[3] / B) * rstd_b
        assert np.allclose(da0, wa[sl], rtol=1e-8, atol=1e-14) and np.allclose(db0, wb[sl], rtol=1e-8, atol=1e-14)

        # Q10: delta is clipped by the GLOBAL min of t -> one MIN all-reduce of an int
        tmin = torch.tensor([int(t[sl].min())]); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        assert int(tmin) == int(t.min())
        q.put((rank, "ok"))
    except Exception as e:                                            # surface the failure in the parent
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_batch_sharded_decomposition_gloo_world2():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


@pytest.mark.parametrize("force_port", [True, False])
def test_bench_reference_arm_contract(force_port):
    """`bench.py --impl reference` prints ONE JSON line with the contract's keys (tiny workload), through the
    unmodified reference when it is importable here and through the port otherwise / when forced."""
    import json
    import subprocess
    import sys
    env = dict(os.environ)
    env.pop("FDDM_BASELINE", None)
    if force_port:
        env["FDDM_BASELINE"] = "port"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None and d["higher_is_better"] is True
    have_ref = any(p and os.path.exists(os.path.join(p, "fddm", "sched", "diffusion_scheduler.py"))
                   for p in (os.environ.get("FDDM_REF"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")))
    assert d["cpu_baseline"]["kind"] == ("reference" if (have_ref and not force_port) else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "slice" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
