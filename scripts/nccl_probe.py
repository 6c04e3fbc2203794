#!/usr/bin/env python
"""Times the all-reduces the sharded step issues (torchrun, NCCL): device time per call, max over ranks."""
import os
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cases = {"scalar f32": torch.zeros(1, device=dev), "cov 768x768 f32 (2.4 MB)": torch.zeros(768 * 768, device=dev),
         "stats 4xTxD f64 (6.3 MB)": torch.zeros(4 * 256 * 768, dtype=torch.float64, device=dev),
         "stats 4xTxD f32 (3.1 MB)": torch.zeros(4 * 256 * 768, device=dev),
         "bn 2xTxD f32 (1.6 MB)": torch.zeros(2 * 256 * 768, device=dev)}
for name, x in cases.items():
    for _ in range(5):
        dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 50], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"allreduce {name:28s} world {world}: {float(ms) * 1e3:7.1f} us")
dist.destroy_process_group()
