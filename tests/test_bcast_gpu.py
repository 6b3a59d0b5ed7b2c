"""world_size-2 test of msfl_bcast_submap through the C ABI (needs 2 GPUs: NCCL refuses two ranks on one device).
Rank 0 owns the submap; rank 1 ADOPTS the broadcast index (no submap_build) and must produce bit-identical poses to
an engine that built the index locally, and poses equal to the oracle's within the parity tolerance."""
import os
import socket

import numpy as np
import pytest

from conftest import make_map_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    from msf_loam_b200 import Engine, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # only carries the 128-byte NCCL id
    case = make_map_case()
    qs = case["queries"]
    eng = Engine(device=rank)
    if rank == 0:
        eng.set_submap(case["map_corner"], case["map_surf"])
    comm = sharding.make_submap_comm(eng)
    for _ in range(2):  # a second map version re-uses the buffers
        eng.bcast_submap(comm, root=0)
    poses = np.stack([eng.scan2map(q["corner"], q["surf"], q["init"], want_stats=False)[1] for q in qs])
    # rank 1: the same scans against a locally built index
    local = Engine(device=rank)
    local.set_submap(case["map_corner"], case["map_surf"])
    poses_local = np.stack([local.scan2map(q["corner"], q["surf"], q["init"], want_stats=False)[1] for q in qs])
    ret[rank] = (poses, poses_local)
    eng.sync()
    eng.nccl_comm_destroy(comm)
    eng.close()
    local.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_bcast_submap_adopts_index_two_ranks():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import oracle as O
    from msf_loam_b200 import synth as S
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    case = make_map_case()
    P = O.default_params()
    for rank in (0, 1):
        poses, poses_local = ret[rank]
        assert np.array_equal(poses, poses_local), f"rank {rank}: adopted index and local index disagree"
    assert np.array_equal(ret[0][0], ret[1][0])
    for q, pose in zip(case["queries"], ret[1][0]):
        ref, _, _ = O.scan2map(P, case["map_corner"], case["map_surf"], q["corner"], q["surf"], q["init"])
        dt, dr = S.pose_error(pose, ref)
        assert dt <= 1e-8 and dr <= 1e-8  # parity bound of north_star is 1e-4; the engine is far inside it


@pytest.mark.gpu
def test_bcast_submap_argument_errors():
    from msf_loam_b200 import Engine, MsflError
    eng = Engine(device=0)
    with pytest.raises(MsflError):
        eng.bcast_submap(0, root=0)  # null communicator
    eng.close()
