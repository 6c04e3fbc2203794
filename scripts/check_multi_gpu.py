#!/usr/bin/env python
"""Multi-GPU correctness check (launch with torchrun, one rank per GPU, NCCL): the batch-sharded
kl_term / lfd_loss (values AND gradients) equal the single-process evaluation on the whole batch.
The check itself is bench.shard_check (bench.py runs it before timing whenever WORLD_SIZE > 1);
this script runs it for per-rank batches 8 (small-batch L_fd path), 32 and 40 (tb-major path, with batch
padding) and exits non-zero on failure."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "fddm-asr_b200")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import fddm_b200 as fb
import bench


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    modes = ["nccl"]
    if fb.symmetric_exchange_available(dist.group.WORLD):        # the library's own all-reduce kernels
        modes += ["p2p", "nvls"]
    elif rank == 0:
        print("multi-gpu check: symmetric memory not available here, only the NCCL exchange is checked", flush=True)
    for collective in modes:
        for per_rank_b in (8, 32, 40):
            res = bench.shard_check(fb, dev, dist.group.WORLD, world, rank, per_rank_b=per_rank_b, collective=collective)
            ok = ok and res["ok"]
            if rank == 0:
                print("multi-gpu check", "OK" if res["ok"] else "FAILED", "world", world, "exchange", collective,
                      json.dumps(res), flush=True)
    bench.teardown(world, dev, code=0 if ok else 1)


if __name__ == "__main__":
    main()
