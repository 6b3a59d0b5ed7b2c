"""The reference-side adapter (msf_loam_b200/adapter/gpu_scan_matchers.h) goes through a compiler and -- on a GPU --
through the engine: it is compiled against the minimal stand-in declarations of tests/adapter_stubs/ (PCL / Eigen /
Ceres / glog / the reference headers are not installed here), linked with libmsfl.so, and driven like
LaserMapping::MatchScan2Map (laser_mapping.cc:304-311) and LaserOdometry::AddLaserScan (laser_odometry.cc:75)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import oracle as O
from msf_loam_b200 import synth as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from msf_loam_b200 import _lib
    _lib.load_library()
    exe = str(tmp_path / "adapter_driver")
    cmd = [shutil.which("g++") or "g++", "-std=c++14", "-O1", "-Wall", "-Werror",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "msf_loam_b200", "adapter"),
           "-I", os.path.join(ROOT, "tests", "adapter_stubs"), os.path.join(ROOT, "tests", "c_abi", "adapter_driver.cc"),
           "-L", os.path.join(ROOT, "msf_loam_b200"), "-lmsfl", "-Wl,-rpath," + os.path.join(ROOT, "msf_loam_b200"), "-o", exe]
    subprocess.run(cmd, check=True)
    return exe


def test_adapter_compiles_as_cxx14_against_stub_declarations(tmp_path):
    """-std=c++14 (the reference's CMAKE_CXX_STANDARD, CMakeLists.txt:5), -Wall -Werror."""
    assert os.path.exists(_build(tmp_path))


def _odometry_pair(case):
    P = O.default_params()
    f0, f1 = case["queries"][0]["features"], case["queries"][1]["features"]
    full0, full1 = f0["full"], f1["full"]
    return (full0[f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]], full0[f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]],
            full1[f1["idx_sharp"]], full1[f1["idx_flat"]])


@pytest.mark.gpu
def test_adapter_poses_equal_the_python_binding(tmp_path, vlp16_case):
    from msf_loam_b200 import Engine, default_params, to_pcl
    exe = _build(tmp_path)
    case = vlp16_case
    q = case["queries"][0]
    lc, rlc, ls, rls, cs, cf = _odometry_pair(case)
    arrays = [case["map_corner"], case["map_surf"], q["corner"], q["surf"], lc, ls, cs, cf]
    path = str(tmp_path / "case.bin")
    with open(path, "wb") as f:
        f.write(struct.pack("8i", *[a.shape[0] for a in arrays]))
        for a in arrays:
            f.write(np.ascontiguousarray(a, dtype=np.float32).tobytes())
        f.write(rlc.astype(np.float32).tobytes())
        f.write(rls.astype(np.float32).tobytes())
        f.write(np.asarray(q["init"], dtype=np.float64).tobytes())
        f.write(S.pose_identity().astype(np.float64).tobytes())
    r = subprocess.run([exe, path], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stdout + r.stderr
    got = {l.split()[0]: np.array(l.split()[1:], dtype=np.float64) for l in r.stdout.splitlines() if l[:3] in ("MAP", "ODO", "DSK", "FRM")}
    eng = Engine(default_params(lm_cluster=16))  # the adapter's engine configuration
    eng.set_submap(case["map_corner"], case["map_surf"])
    _, pose_map, _ = eng.scan2map(to_pcl(q["corner"]), to_pcl(q["surf"]), q["init"])
    rc, pose_odo, _ = eng.scan2scan(to_pcl(lc, rlc), to_pcl(ls, rls), to_pcl(cs, np.zeros(len(cs))), to_pcl(cf, np.zeros(len(cf))),
                                    S.pose_identity())
    assert got["MAP"][0] == 1 and np.array_equal(got["MAP"][1:], pose_map)
    assert got["ODO"][0] == (1 if rc == 0 else 0) and np.array_equal(got["ODO"][1:], pose_odo)
    # IMU-initialised branch with an identity preintegration and zero velocity / gravity: the deskew factors reduce to
    # the plain ones, so the pose agrees with the LiDAR-only solve to solver precision
    dt, dr = S.pose_error(got["DSK"][1:], pose_map)
    assert got["DSK"][0] == 1 and dt < 1e-6 and dr < 1e-6
    # and with the oracle (the parity bound of north_star is 1e-4)
    ref, _, _ = O.scan2map(O.default_params(), case["map_corner"], case["map_surf"], q["corner"], q["surf"], q["init"])
    dt, dr = S.pose_error(got["MAP"][1:], ref)
    assert dt <= 1e-7 and dr <= 1e-7
    # GpuMappingFrame: the first frame only fills the maps, the second matches against them.  The scan's feature clouds go
    # through the maps' VoxelGrid (0.2 / 0.4) inside the call and the map clouds are re-filtered per 3 m cell on insert, so
    # the pose is close to -- not bitwise -- the MatchScan2Map one above
    dt, dr = S.pose_error(got["FRM"][1:], q["gt"])
    assert got["FRM"][0] == 1 and dt < 0.05 and dr < 0.01
    dt, dr = S.pose_error(got["FRM"][1:], pose_map)
    assert dt < 0.02 and dr < 0.005
    eng.close()
