#!/usr/bin/env python
"""Probe (torchrun, N >= 2): is symmetric memory (peer mappings, multicast / NVLS mapping) available on this box,
is `fddm_xgpu_allreduce` correct on it (both algorithms, fp32 and fp64, eager and replayed in a CUDA graph), and
how long does it take next to ncclAllReduce for the three L_fd exchange sizes at c5shard (T=256, D=768)?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/nvls_probe.py
Rank 0 prints one JSON line."""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fddm-asr_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist


def timed(fn, iters=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 2)                                    # us per call, max over ranks


def graphed(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    return g, reps


MAX_CTAS = int(os.environ.get('XGPU_MAX_CTAS', '0'))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    code = 0
    try:
        import torch.distributed._symmetric_memory as symm
        import fddm_b200 as fb
        L = fb._lib
        gname = dist.group.WORLD.group_name
        T, D = 256, 768
        sizes = {"tiny_f32": (64, torch.float32), "stats_f64": (4 * T * D, torch.float64),
                 "cov_f32": (D * D + 4, torch.float32), "bn_f32": (2 * T * D, torch.float32)}
        out["signal_pad_size"] = int(symm.get_signal_pad_size())
        out["pad_bytes_needed"] = int(L.lib.fddm_xgpu_signal_pad_bytes())
        for name, (n, dt) in sizes.items():
            buf = symm.empty(n, dtype=dt, device=dev)
            hdl = symm.rendezvous(buf, gname)
            mc = int(hdl.multicast_ptr)
            out.setdefault("multicast", bool(mc))
            pads = int(hdl.signal_pad_ptrs_dev)
            bufs = int(hdl.buffer_ptrs_dev)
            eb = 8 if dt == torch.float64 else 4

            def mine(algo=L.XGPU_P2P):
                L.check(L.lib.fddm_xgpu_allreduce(bufs, mc or None, pads, rank, world, eb, n, algo, MAX_CTAS, L.stream_ptr(dev)),
                        "xgpu_allreduce")

            def mine_nvls():
                mine(L.XGPU_NVLS)

            gen = torch.Generator(device=dev).manual_seed(1234 + rank)
            src = torch.randn(n, generator=gen, device=dev, dtype=dt)
            want = src.clone()
            dist.all_reduce(want)
            buf.copy_(src)
            torch.cuda.synchronize(); dist.barrier()
            mine()
            torch.cuda.synchronize()
            err = float((buf - want).abs().max() / want.abs().max())
            same = buf.clone()
            dist.broadcast(same, 0)
            out[name] = {"n": n, "rel_err_vs_nccl": err, "bitwise_same_on_all_ranks": bool(torch.equal(same, buf))}
            if not err < (1e-12 if eb == 8 else 1e-5):
                code = 1
            # CUDA-graph replay: 20 all-reduces of a buffer refilled inside the graph
            def refill_and_reduce():
                buf.copy_(src)
                mine()
            g, reps = graphed(refill_and_reduce)
            g.replay(); g.replay()
            torch.cuda.synchronize()
            out[name]["graph_replay_rel_err"] = float((buf - want).abs().max() / want.abs().max())
            if not out[name]["graph_replay_rel_err"] < (1e-12 if eb == 8 else 1e-5):
                code = 1
            out[name]["us_p2p_eager"] = timed(mine)
            g2, reps = graphed(mine)
            out[name]["us_p2p_graph"] = round(timed(g2.replay, iters=10, warm=2) / reps, 2)
            if mc:
                buf.copy_(src)
                torch.cuda.synchronize(); dist.barrier()
                mine_nvls()
                torch.cuda.synchronize()
                out[name]["nvls_rel_err_vs_nccl"] = float((buf - want).abs().max() / want.abs().max())
                out[name]["us_nvls_eager"] = timed(mine_nvls)
                g4, reps = graphed(mine_nvls)
                out[name]["us_nvls_graph"] = round(timed(g4.replay, iters=10, warm=2) / reps, 2)
                del g4
            nc = src.clone()
            out[name]["us_nccl_eager"] = timed(lambda: dist.all_reduce(nc))
            g3, reps = graphed(lambda: dist.all_reduce(nc))
            out[name]["us_nccl_graph"] = round(timed(g3.replay, iters=10, warm=2) / reps, 2)
            if dt == torch.float32:
                for opname in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
                    try:
                        op = getattr(torch.ops.symm_mem, opname)
                        out[name]["us_torch_" + opname] = timed(lambda: op(buf, "sum", gname))
                    except Exception as e:                      # library op not available for this build / size
                        out[name]["us_torch_" + opname] = f"{type(e).__name__}: {str(e)[:80]}"
            del g, g2, g3
    except Exception as e:
        out["error"] = f"{type(e).__name__}: {e}"
        out["trace"] = traceback.format_exc()[-800:]
        code = 2
    if rank == 0:
        print(json.dumps(out), flush=True)
    torch.cuda.synchronize()
    try:
        dist.barrier()
        dist.destroy_process_group()
    finally:
        os._exit(code)


if __name__ == "__main__":
    main()
