"""Multi-GPU plumbing for the scan-matching path: independent scans are sharded over
ranks (one process per GPU), the shared submap is broadcast once per map version, poses are gathered.
There is no per-iteration exchange -- each scan-to-map solve depends only on (scan, submap, guess).
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_scans: int, rank: int, world: int) -> np.ndarray:
    """scan i -> rank i mod world (round-robin keeps ragged batches balanced)."""
    return np.arange(rank, n_scans, world, dtype=np.int64)


def broadcast_submap(map_corner, map_surf, src: int = 0, device=None):
    """Broadcasts the two float4 submap arrays from `src`.  On `src` the inputs are (n,4) float32
    tensors/arrays; elsewhere they may be None.  Returns (corner, surf) tensors on `device`."""
    rank = dist.get_rank()
    if rank == src:
        tc = torch.as_tensor(map_corner, dtype=torch.float32).reshape(-1, 4)
        ts = torch.as_tensor(map_surf, dtype=torch.float32).reshape(-1, 4)
        if device is not None:
            tc, ts = tc.to(device), ts.to(device)
        sizes = torch.tensor([tc.shape[0], ts.shape[0]], dtype=torch.int64, device=tc.device)
    else:
        sizes = torch.zeros(2, dtype=torch.int64, device=device)
    dist.broadcast(sizes, src)
    if rank != src:
        tc = torch.empty((int(sizes[0]), 4), dtype=torch.float32, device=device)
        ts = torch.empty((int(sizes[1]), 4), dtype=torch.float32, device=device)
    dist.broadcast(tc, src)
    dist.broadcast(ts, src)
    return tc, ts


def gather_poses(local_poses, n_scans: int, device=None) -> np.ndarray:
    """All-gathers the per-rank (n_local,7) poses back into scan order (n_scans,7)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    n_max = (n_scans + world - 1) // world
    buf = torch.zeros((n_max, 7), dtype=torch.float64, device=device)
    lp = torch.as_tensor(np.asarray(local_poses), dtype=torch.float64).reshape(-1, 7)
    buf[: lp.shape[0]] = lp.to(buf.device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    poses = np.zeros((n_scans, 7))
    for r in range(world):
        idx = shard_indices(n_scans, r, world)
        poses[idx] = out[r][: len(idx)].cpu().numpy()
    return poses


def nccl_unique_id(engine) -> bytes:
    """Rank 0: ncclGetUniqueId through the C ABI (msfl_nccl_get_unique_id)."""
    import ctypes as C
    buf = (C.c_ubyte * 128)()
    engine._check(engine.lib.msfl_nccl_get_unique_id(buf))
    return bytes(buf)


def make_submap_comm(engine, device=None) -> int:
    """An NCCL communicator for msfl_bcast_submap with the ranks of the default torch.distributed group: rank 0's
    unique id travels over the existing process group (any backend), then every rank calls msfl_nccl_comm_init.
    This is what a C++ host does with its own out-of-band channel."""
    rank, world = dist.get_rank(), dist.get_world_size()
    ident = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        ident = torch.frombuffer(bytearray(nccl_unique_id(engine)), dtype=torch.uint8).to(ident.device)
    dist.broadcast(ident, 0)
    return engine.nccl_comm_init(bytes(ident.cpu().numpy().tobytes()), world, rank)
