"""How fast can this box feed its GPUs?  Every rank copies a pinned 256 MiB buffer to its GPU 20 times, all ranks at once;
prints per-rank and aggregate host->device GB/s.  (torchrun --nproc-per-node N tools/dev_h2d_cap.py)"""
import os
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbs = torch.tensor([20 * n / dt / 1e9], dtype=torch.float64, device=dev)
allg = [torch.zeros_like(gbs) for _ in range(world)]
if world > 1:
    dist.all_gather(allg, gbs)
else:
    allg = [gbs]
if rank == 0:
    v = [round(float(x), 1) for x in allg]
    print(f"N={world} per-rank H2D GB/s {v} aggregate {sum(v):.1f}", flush=True)
if world > 1:
    dist.destroy_process_group()
