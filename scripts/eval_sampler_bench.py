#!/usr/bin/env python
"""the reference's evaluation decodes utterance by utterance -- it builds one B=1
`DiffusionJumpySampler` per utterance (models/evaluate.py:163-176).  This script times that call pattern against the
batched `sample_utterances` on one GPU, with a dummy decoder that returns pre-generated logits (the decoder is out
of scope), greedy "map" decoding, T_infer=20, r=5 (4 jumps):

    python scripts/eval_sampler_bench.py [--B 32] [--L 128] [--V 8000] [--mode exact] [--iters 10]

Prints one JSON line: wall-clock ms per batch of B utterances (host sync on both sides -- host overhead is the
point here) and tokens*V/s for (a) a fresh B=1 sampler per utterance, (b) one B=1 sampler replaying its CUDA graph
per utterance, (c) `sample_utterances` eager, (d) `sample_utterances` with the chain in one CUDA graph."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fddm-asr_b200")):
    sys.path.insert(0, p)
import torch
import fddm_b200 as fb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--L", type=int, default=128)
    ap.add_argument("--V", type=int, default=8000)
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"])
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, L, V = a.B, a.L, a.V
    g = torch.Generator(device=dev).manual_seed(1337)
    logits = torch.randn(4, B, L, V, generator=g, device=dev) * 4
    cond = torch.zeros(B, 1, 1, device=dev)
    cond[:, 0, 0] = torch.arange(B, device=dev)                 # the utterance index travels in cond
    sch = fb.DiscreteDiffusionScheduler(K=V, T=200, device=dev)

    flat = logits.view(4 * B, L, V)

    def dec(x, t, c):                                           # capturable (no host read); cost ~ rows returned,
        idx = (t // 5 - 1).clamp(0, 3) * B + c[:, 0, 0].long()  # so the B=1 calls are not charged for the whole batch
        return torch.index_select(flat, 0, idx)

    mk = lambda: fb.DiffusionJumpySampler(sch, dec, K=V, T_train=200, T_infer=20, r=5, greedy=True,
                                          posterior_mode="map", sampling_mode=a.mode, device=dev)
    x_T = torch.randint(0, V, (B, L), generator=g, device=dev)

    def per_utterance_fresh():
        return torch.cat([mk().sample(cond[b:b + 1], L, x_init=x_T[b:b + 1], return_p=False)[0] for b in range(B)])

    one = mk().enable_cuda_graph()

    def per_utterance_graph():
        return torch.cat([one.sample(cond[b:b + 1], L, x_init=x_T[b:b + 1], return_p=False)[0].clone()
                          for b in range(B)])

    batched = mk()
    batched_g = mk().enable_cuda_graph()
    cases = {
        "per_utterance_fresh_sampler": per_utterance_fresh,
        "per_utterance_graph_replay": per_utterance_graph,
        "sample_utterances": lambda: batched.sample_utterances(cond, L, x_init=x_T)[0].view(B, L),
        "sample_utterances_graph": lambda: batched_g.sample_utterances(cond, L, x_init=x_T)[0].view(B, L).clone(),
    }
    out = {"workload": f"eval decode B={B} L={L} V={V} greedy map {a.mode} 4 jumps", "iters": a.iters, "cases": {}}
    want = None
    for name, fn in cases.items():
        for _ in range(3):
            ids = fn()
        torch.cuda.synchronize()
        if want is None:
            want = ids.clone()
        same = bool(torch.equal(ids, want))
        t0 = time.perf_counter()
        for _ in range(a.iters):
            fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / a.iters
        out["cases"][name] = {"ms_per_batch": round(ms, 3), "G_tokV_per_s": round(4 * B * L * V / ms / 1e6, 2),
                              "ids_equal_to_first_case": same}
    base = out["cases"]["per_utterance_fresh_sampler"]["ms_per_batch"]
    for c in out["cases"].values():
        c["speedup_vs_per_utterance"] = round(base / c["ms_per_batch"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
