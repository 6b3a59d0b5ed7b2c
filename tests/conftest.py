import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# ------------------------------------------------------------------ shared synthetic workloads
import oracle as O  # noqa: E402  (tests are allowed to use the oracle)
from msf_loam_b200 import synth as S  # noqa: E402

_CACHE = {}


def make_map_case(sensor="vlp16", scene_kind="room40", n_map_scans=5, seed0=100, sigma=0.01):
    """BASELINE config-2-shaped case built with the ORACLE's extraction (tests only):
    submap = features of scans 0..n-1 at GT poses, VoxelGrid 0.2/0.4; query = scan n."""
    key = (sensor, scene_kind, n_map_scans, seed0, sigma)
    if key in _CACHE:
        return _CACHE[key]
    P = O.default_params()
    scene = S.make_scene(scene_kind)
    traj = S.trajectory(n_map_scans + 3)
    mc, ms = [], []
    for k in range(n_map_scans):
        xyzi, ring = S.raycast_scan(scene, sensor, traj[k], seed=seed0 + k, sigma=sigma)
        f = O.extract_features(P, xyzi, ring, S.pose_identity())
        mc.append(S.transform_cloud(traj[k], f["full"][f["idx_less_sharp"]]))
        ms.append(S.transform_cloud(traj[k], f["full"][f["idx_less_flat"]]))
    map_corner = O.voxel_grid(np.concatenate(mc), 0.2)
    map_surf = O.voxel_grid(np.concatenate(ms), 0.4)
    queries = []
    for k in range(n_map_scans, n_map_scans + 3):
        xyzi, ring = S.raycast_scan(scene, sensor, traj[k], seed=seed0 + k, sigma=sigma)
        f = O.extract_features(P, xyzi, ring, S.pose_identity())
        sc = O.voxel_grid(f["full"][f["idx_less_sharp"]], 0.2)
        ss = O.voxel_grid(f["full"][f["idx_less_flat"]], 0.4)
        rng = np.random.default_rng(seed0 + 1000 + k)
        queries.append({"corner": sc, "surf": ss, "gt": traj[k], "init": S.perturb_pose(traj[k], rng),
                        "raw": (xyzi, ring), "features": f})
    case = {"map_corner": map_corner, "map_surf": map_surf, "queries": queries, "traj": traj, "scene": scene}
    _CACHE[key] = case
    return case


@pytest.fixture(scope="session")
def vlp16_case():
    return make_map_case()
