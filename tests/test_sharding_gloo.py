"""N>1 host logic on CPU: world_size-2 gloo run of scan sharding + submap broadcast + pose gather.
The per-rank matcher here is the CPU oracle (no GPU in this container); the GPU box runs the same
plumbing with the CUDA engine in bench.py."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from msf_loam_b200 import sharding
from conftest import make_map_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_scans, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = make_map_case()
    qs = case["queries"]
    # only rank 0 "owns" the submap
    tc, ts = sharding.broadcast_submap(case["map_corner"] if rank == 0 else None,
                                       case["map_surf"] if rank == 0 else None, src=0)
    assert tc.shape[0] == case["map_corner"].shape[0] and ts.shape[0] == case["map_surf"].shape[0]
    mine = sharding.shard_indices(n_scans, rank, world)
    P = O.default_params()
    local = [O.scan2map(P, tc.numpy(), ts.numpy(), qs[i % len(qs)]["corner"], qs[i % len(qs)]["surf"],
                        qs[i % len(qs)]["init"])[0] for i in mine]
    poses = sharding.gather_poses(np.array(local).reshape(-1, 7), n_scans)
    if rank == 0:
        ret["poses"] = poses
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    n_scans = 5  # ragged: 3 + 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n_scans, ret), nprocs=2, join=True)
    case = make_map_case()
    qs = case["queries"]
    P = O.default_params()
    ref = np.stack([O.scan2map(P, case["map_corner"], case["map_surf"], qs[i % 3]["corner"], qs[i % 3]["surf"],
                               qs[i % 3]["init"])[0] for i in range(n_scans)])
    assert np.array_equal(ret["poses"], ref)


def test_shard_indices_cover_everything_once():
    for n in (0, 1, 7, 64):
        for w in (1, 2, 4, 8):
            allidx = np.concatenate([sharding.shard_indices(n, r, w) for r in range(w)])
            assert sorted(allidx.tolist()) == list(range(n))
