"""The REFERENCE's own scan registration as the checker (SURVEY.md section 8 rows a-1 .. a-4).

oracle/_ref/libmsfl_ref.so contains src/msf_loam_node.cc compiled whole and unmodified (oracle/ref_extract_shim.cc
#includes it from the reference checkout; ROS / rosbag / protobuf / PCL stood in by oracle/ref_stubs/), and the harness
runs its RealHandleLaserCloudMessage (msf_loam_node.cc:160-378) on the same raw clouds as the oracle and the CUDA path.
What the reference hands to LaserOdometry::AddLaserScan -- the registered full cloud and the four feature clouds, after
the extrinsic -- is compared: bit-exact xyz / ring / membership / order; the relative time (intensity) bit-exact on the
CPU (both sides call glibc's atan2f) and to 1.6e-8 s on the GPU (see tests/test_features_gpu.py).
"""
import numpy as np
import pytest

import oracle as O
from oracle import ref as R
from msf_loam_b200 import synth as S

pytestmark = pytest.mark.skipif(not R.available(), reason="no reference checkout and no prebuilt oracle/_ref library")

T_EXT = np.concatenate([[0.1, -0.2, 0.3], S.rotvec_to_quat(np.array([0.02, -0.01, 0.3]))])
LISTS = (("sharp", "idx_sharp"), ("less_sharp", "idx_less_sharp"), ("flat", "idx_flat"), ("less_flat", "idx_less_flat"))


def _dirty_ring_major_scan():
    """ring-major input, NaN / inf / too-close points, one ring too short to be used (msf_loam_node.cc:252)."""
    xyzi, ring = S.raycast_scan(S.make_scene(), "vlp16", S.trajectory(1)[0], seed=7)
    order = np.argsort(ring, kind="stable")
    xyzi, ring = xyzi[order].copy(), ring[order].copy()
    bad = np.random.default_rng(3).choice(len(xyzi), 300, replace=False)
    xyzi[bad[:100], 0] = np.nan
    xyzi[bad[100:200], :3] *= 1e-3
    xyzi[bad[200:], 2] = np.inf
    keep = ~((ring == 4) & (np.arange(len(ring)) % 200 != 0))
    return xyzi[keep], ring[keep]


def _cases():
    sc = S.make_scene("room80")
    traj = S.trajectory(3)
    yield "vlp16", S.raycast_scan(S.make_scene(), "vlp16", traj[1], seed=41), T_EXT
    yield "hdl64", S.raycast_scan(sc, "hdl64", traj[0], seed=42), T_EXT
    yield "os1-128", S.raycast_scan(sc, "os1-128", traj[2], seed=43), None
    yield "dirty", _dirty_ring_major_scan(), None


def _compare(f, r, time_tol):
    """f: dict with `full`, `ring` and index lists (oracle / CUDA); r: the reference's clouds."""
    assert f["full"].shape == r["full"].shape and np.array_equal(f["ring"], r["ring"])
    assert np.array_equal(f["full"][:, :3], r["full"][:, :3])
    assert np.abs(f["full"][:, 3] - r["full"][:, 3]).max() <= time_tol
    for cloud, idx in LISTS:
        got = f["full"][f[idx]]
        assert got.shape == r[cloud].shape, cloud
        assert np.array_equal(got[:, :3], r[cloud][:, :3]), cloud
        assert np.abs(got[:, 3] - r[cloud][:, 3]).max() <= time_tol, cloud


def test_oracle_extraction_equals_the_reference_registration():
    P = O.default_params()
    for name, (xyzi, ring), T in _cases():
        r = R.extract_features(xyzi, ring, T)
        f = O.extract_features(P, xyzi, ring, T)
        assert r["sharp"].shape[0] > 50 and r["flat"].shape[0] > 100 and r["less_flat"].shape[0] > 10000, name
        _compare(f, r, 0.0)


def test_oracle_extraction_equals_the_reference_registration_other_min_range():
    """minimum_range is a node parameter (msf_loam_node.cc:434): 8 m removes the nearest floor rings and walls."""
    xyzi, ring = S.raycast_scan(S.make_scene(), "vlp16", S.trajectory(2)[1], seed=11)
    P = O.default_params(min_range=8.0)
    r = R.extract_features(xyzi, ring, T_EXT, min_range=8.0)
    f = O.extract_features(P, xyzi, ring, T_EXT)
    assert r["full"].shape[0] < xyzi.shape[0] - 100
    _compare(f, r, 0.0)


def test_reference_less_flat_cloud_is_not_downsampled():
    """Quirk Q1: VoxelGridWrapper copies the filter's INPUT indices (msf_loam_node.cc:122-125), so the less-flat cloud is
    every FLAT / UNKNOWN point of the used sectors, far more than a 0.2 m voxel grid would leave."""
    xyzi, ring = S.raycast_scan(S.make_scene(), "vlp16", S.trajectory(1)[0], seed=5)
    r = R.extract_features(xyzi, ring, None)
    assert r["less_flat"].shape[0] > 0.7 * r["full"].shape[0]
    assert O.voxel_grid(r["less_flat"], 0.2).shape[0] < 0.5 * r["less_flat"].shape[0]


@pytest.mark.gpu
def test_cuda_extraction_equals_the_reference_registration():
    from msf_loam_b200 import Engine
    e = Engine()
    try:
        for name, (xyzi, ring), T in _cases():
            r = R.extract_features(xyzi, ring, T)
            g = e.extract_features(xyzi, ring, T)
            _compare(g, r, 1.6e-8)
    finally:
        e.close()
