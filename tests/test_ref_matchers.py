"""The REFERENCE's own matchers as the checker (SURVEY.md section 8 rows a-5, a-6, a-7, a-9 call sites, a-10, f-3).

oracle/_ref/libmsfl_ref.so holds odometry_scan_matcher.cc, mapping_scan_matcher.cc, scan_matcher.cc, lidar_factor.cc,
pose_local_parameterization.cc and scan_undistortion.cc compiled UNMODIFIED from the reference checkout against the
stand-in third-party headers of oracle/ref_stubs/ (oracle/ref_shim.cc: what is the reference's -- every association
loop, gate, fit expression, factor, parameter block, iteration cap -- and what is stood in -- exact k-NN, 3x3 eigen /
5x3 QR, Eigen's small fixed-size algebra, the Ceres trust-region loop).  CPU tests pin the oracle's restatement to it;
GPU tests compare the CUDA path through the C ABI with it directly.
"""
import numpy as np
import pytest

import oracle as O
from oracle import ref as R
from msf_loam_b200 import synth as S
from test_deskew import GRAV, VEL, make_preintegration

pytestmark = pytest.mark.skipif(not R.available(), reason="no reference checkout and no prebuilt oracle/_ref library")


def _scan_pair(seed0=100):
    P = O.default_params()
    sc = S.make_scene()
    traj = S.trajectory(3)
    f = [O.extract_features(P, *S.raycast_scan(sc, "vlp16", traj[k], seed=seed0 + k), None) for k in range(2)]
    lc, lcr = f[0]["full"][f[0]["idx_less_sharp"]], f[0]["ring"][f[0]["idx_less_sharp"]]
    ls, lsr = f[0]["full"][f[0]["idx_less_flat"]], f[0]["ring"][f[0]["idx_less_flat"]]
    cs, cf = f[1]["full"][f[1]["idx_sharp"]], f[1]["full"][f[1]["idx_flat"]]
    return P, (lc, lcr, ls, lsr, cs, cf), S.pose_mul(S.pose_inv(traj[0]), traj[1])


def _same_trace(lm_ref, lm_oracle, rtol=1e-9):
    assert lm_ref["n_attempts"] == lm_oracle["n_attempts"] and lm_ref["termination"] == lm_oracle["termination"]
    for a, b in zip(lm_ref["iters"], lm_oracle["iters"]):
        assert a["accepted"] == b["accepted"] and a["valid"] == b["valid"]
        assert abs(a["cost"] - b["cost"]) <= rtol * abs(b["cost"])
        assert abs(a["radius"] - b["radius"]) <= 1e-6 * abs(b["radius"])


def test_transform_point_is_the_reference_transform():
    rng = np.random.default_rng(3)
    for _ in range(20):
        pose = np.concatenate([rng.normal(scale=30, size=3), S.rotvec_to_quat(rng.normal(scale=0.8, size=3))])
        xyz = rng.normal(scale=40, size=(500, 3)).astype(np.float32)
        assert np.array_equal(R.transform_point(pose, xyz), O.transform_points_f(pose, xyz))


def test_get_delta_qp_is_the_reference_interpolation():
    import ctypes as C
    t, dq, dp = make_preintegration()
    dk = O.Deskew(O._ptr(t, C.c_double), O._ptr(np.ascontiguousarray(dq), C.c_double), O._ptr(np.ascontiguousarray(dp), C.c_double),
                  t.shape[0], (C.c_double * 3)(0, 0, 0), (C.c_double * 3)(0, 0, 0))
    rng = np.random.default_rng(4)
    for dt in list(rng.uniform(t[0], t[-1] * 0.999, size=300)) + [float(t[0]), float(t[7]), float(np.float32(0.0421))]:
        q, p = np.zeros(4), np.zeros(3)
        assert O.lib().msflo_get_delta_qp(C.byref(dk), C.c_double(dt), O._ptr(q, C.c_double), O._ptr(p, C.c_double)) == 0
        q_ref, p_ref = R.get_delta_qp(t, dq, dp, dt)
        assert np.array_equal(q, q_ref) and np.array_equal(p, p_ref)


def test_oracle_scan2map_equals_the_reference_matcher(vlp16_case):
    P = O.default_params()
    c = vlp16_case
    for q in c["queries"]:
        R.reset_logs(record_knn=True)
        ok, x_ref = R.scan2map(c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
        solves = R.solves()
        k_of, idx, d2 = R.knn_log()
        x, logs, counts = O.scan2map(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
        assert ok is True and len(solves) == 2 and all(s["supported"] for s in solves)
        # correspondences the reference created in each outer iteration (mapping_scan_matcher.cc:172,241)
        assert [(s["n_edge"], s["n_plane"]) for s in solves] == [tuple(r) for r in counts]
        dt, dr = S.pose_error(x, x_ref)
        assert dt < 1e-9 and dr < 1e-9, (dt, dr)
        for s, lg in zip(solves, logs):
            _same_trace(s["lm"], lg)
        # the reference's own sequence of 5-NN searches (outer iteration 0: corner queries, then surf queries) against
        # the oracle's association at the same pose
        nq = q["corner"].shape[0] + q["surf"].shape[0]
        assert k_of.shape[0] == 2 * nq and np.all(k_of == 5)
        _, _, _, kidx = O.associate_map(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
        first = idx.reshape(-1, 5)[:nq]
        gate = d2.reshape(-1, 5)[:nq, 4] < 1.0
        assert np.array_equal(gate, kidx[:, 0] >= 0)
        assert np.array_equal(first[gate], kidx[gate])


def test_oracle_scan2scan_equals_the_reference_matcher():
    P, (lc, lcr, ls, lsr, cs, cf), gt = _scan_pair()
    for init in (S.pose_identity(), gt):
        R.reset_logs()
        ok, x_ref = R.scan2scan(lc, lcr, ls, lsr, cs, cf, init)
        solves = R.solves()
        rc, x, logs, counts, _ = O.scan2scan(P, lc, lcr, ls, lsr, cs, cf, init)
        assert ok is True and rc == 0 and len(solves) == 2
        assert [(s["n_edge"], s["n_plane"]) for s in solves] == [tuple(r) for r in counts]
        assert counts[0, 0] > 50 and counts[0, 1] > 200
        dt, dr = S.pose_error(x, x_ref)
        assert dt < 1e-9 and dr < 1e-9, (dt, dr)
        for s, lg in zip(solves, logs):
            _same_trace(s["lm"], lg)
    # fewer than 10 correspondences: false, pose untouched, no solve (odometry_scan_matcher.cc:262-267)
    R.reset_logs()
    init = np.array([0.01, 0.02, 0.0, 0, 0, 0, 1.0])
    ok, x_ref = R.scan2scan(lc, lcr, ls, lsr, cs[:3], cf[:4], init)
    rc, x, _, _, _ = O.scan2scan(P, lc, lcr, ls, lsr, cs[:3], cf[:4], init)
    assert ok is False and rc == 1 and R.solves() == []
    assert np.array_equal(x_ref, init) and np.array_equal(x, init)


def test_oracle_scan2scan_equals_the_reference_matcher_hdl64():
    """64 rings: the +-2.5-ring windows of the odometry association (odometry_scan_matcher.cc:96-141, :171-230) see many
    more neighbouring rings than on a VLP-16."""
    P = O.default_params()
    sc = S.make_scene("room80")
    traj = S.trajectory(3)
    f = [O.extract_features(P, *S.raycast_scan(sc, "hdl64", traj[k], seed=70 + k), None) for k in range(2)]
    lc, lcr = f[0]["full"][f[0]["idx_less_sharp"]], f[0]["ring"][f[0]["idx_less_sharp"]]
    ls, lsr = f[0]["full"][f[0]["idx_less_flat"]], f[0]["ring"][f[0]["idx_less_flat"]]
    cs, cf = f[1]["full"][f[1]["idx_sharp"]], f[1]["full"][f[1]["idx_flat"]]
    R.reset_logs()
    ok, x_ref = R.scan2scan(lc, lcr, ls, lsr, cs, cf, S.pose_identity())
    solves = R.solves()
    rc, x, logs, counts, _ = O.scan2scan(P, lc, lcr, ls, lsr, cs, cf, S.pose_identity())
    assert ok and rc == 0 and np.array_equal(x, x_ref)
    assert [(s["n_edge"], s["n_plane"]) for s in solves] == [tuple(r) for r in counts] and counts[0, 1] > 1000
    dt, dr = S.pose_error(x, S.pose_mul(S.pose_inv(traj[0]), traj[1]))
    assert dt < 0.05 and dr < 0.01


def test_oracle_deskew_branch_equals_the_reference_matcher(vlp16_case):
    P = O.default_params()
    c = vlp16_case
    t, dq, dp = make_preintegration()
    for q in c["queries"][:2]:
        R.reset_logs()
        ok, x_ref, v_ref = R.scan2map_deskew(c["map_corner"], c["map_surf"], q["corner"], q["surf"], t, dq, dp, VEL, GRAV, q["init"])
        solves = R.solves()
        rc, x, logs, counts, _ = O.scan2map_deskew(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], t, dq, dp, VEL,
                                                   GRAV, q["init"])
        # three Solve calls: the IMU-only predict (declined: side-car, out of scope) and the two outer iterations
        assert ok is True and rc == 0 and [s["supported"] for s in solves] == [False, True, True]
        assert [(s["n_edge"], s["n_plane"]) for s in solves[1:]] == [tuple(r) for r in counts]
        dt, dr = S.pose_error(x, x_ref)
        assert dt < 1e-9 and dr < 1e-9, (dt, dr)
        assert np.array_equal(v_ref, np.array(VEL))  # the speed-bias block is held constant (mapping_scan_matcher.cc:94)
        for s, lg in zip(solves[1:], logs):
            _same_trace(s["lm"], lg)


def test_oracle_scan2map_equals_the_reference_matcher_hdl64_and_degenerate_inputs(vlp16_case):
    """The 64-ring case (12.6 k queries, 43 k-point submap) and the inputs the reference handles without a solve: a scan
    with no corner features, and an initial guess so far off that no query passes the d5^2 < 1 gate (no residual blocks:
    the pose comes back untouched)."""
    from conftest import make_map_case
    P = O.default_params()
    c = make_map_case("hdl64", "room80", 5, 200)
    q = c["queries"][0]
    R.reset_logs()
    ok, x_ref = R.scan2map(c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
    solves = R.solves()
    x, logs, counts = O.scan2map(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
    assert ok and np.array_equal(x, x_ref)
    assert [(s["n_edge"], s["n_plane"]) for s in solves] == [tuple(r) for r in counts] and counts[0, 1] > 5000
    for s, lg in zip(solves, logs):
        _same_trace(s["lm"], lg)
    c, q = vlp16_case, vlp16_case["queries"][0]
    none = np.zeros((0, 4), np.float32)
    R.reset_logs()
    ok, x_ref = R.scan2map(c["map_corner"], c["map_surf"], none, q["surf"], q["init"])
    x, _, counts = O.scan2map(P, c["map_corner"], c["map_surf"], none, q["surf"], q["init"])
    assert ok and np.array_equal(x, x_ref) and [s["n_edge"] for s in R.solves()] == [0, 0] and list(counts[:, 0]) == [0, 0]
    far = q["init"].copy()
    far[:3] += 500.0
    R.reset_logs()
    ok, x_ref = R.scan2map(c["map_corner"], c["map_surf"], q["corner"], q["surf"], far)
    x, _, counts = O.scan2map(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], far)
    assert ok and np.array_equal(x_ref, far) and np.array_equal(x, far) and not counts.any()
    assert all(s["n_edge"] + s["n_plane"] == 0 for s in R.solves())


# ------------------------------------------------------------------------------------------------ GPU vs the reference
@pytest.mark.gpu
def test_cuda_scan2map_equals_the_reference_matcher(vlp16_case):
    from msf_loam_b200 import Engine
    c = vlp16_case
    e = Engine()
    try:
        e.set_submap(c["map_corner"], c["map_surf"])
        for q in c["queries"]:
            R.reset_logs(record_knn=True)
            ok, x_ref = R.scan2map(c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
            solves = R.solves()
            _, idx, d2 = R.knn_log()
            rc, x, st = e.scan2map(q["corner"], q["surf"], q["init"])
            assert ok is True and rc == 0
            dt, dr = S.pose_error(x, x_ref)
            assert dt < 1e-8 and dr < 1e-8, (dt, dr)  # north_star tolerance 1e-4; measured ~1e-14
            assert st["n_edge"] == [s["n_edge"] for s in solves] and st["n_plane"] == [s["n_plane"] for s in solves]
            assert [l["n_attempts"] for l in st["lm"]] == [s["lm"]["n_attempts"] for s in solves]
            # the CUDA 5-NN against the reference's own searches of outer iteration 0: bit-exact indices
            nq = q["corner"].shape[0] + q["surf"].shape[0]
            knn, _ = e.associate_map(q["corner"], q["surf"], q["init"])
            gate = d2.reshape(-1, 5)[:nq, 4] < 1.0
            assert np.array_equal(gate, knn[:, 0] >= 0)
            assert np.array_equal(idx.reshape(-1, 5)[:nq][gate], knn[gate])
    finally:
        e.close()


@pytest.mark.gpu
def test_cuda_scan2map_equals_the_reference_matcher_hdl64_and_degenerate_inputs(vlp16_case):
    from conftest import make_map_case
    from msf_loam_b200 import Engine
    c = make_map_case("hdl64", "room80", 5, 200)
    e = Engine()
    try:
        e.set_submap(c["map_corner"], c["map_surf"])
        for q in c["queries"][:2]:
            ok, x_ref = R.scan2map(c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
            rc, x, st = e.scan2map(q["corner"], q["surf"], q["init"])
            dt, dr = S.pose_error(x, x_ref)
            assert ok and rc == 0 and dt < 1e-8 and dr < 1e-8, (dt, dr)
        c, q = vlp16_case, vlp16_case["queries"][0]
        e.set_submap(c["map_corner"], c["map_surf"])
        none = np.zeros((0, 4), np.float32)
        ok, x_ref = R.scan2map(c["map_corner"], c["map_surf"], none, q["surf"], q["init"])
        rc, x, st = e.scan2map(none, q["surf"], q["init"])
        dt, dr = S.pose_error(x, x_ref)
        assert rc == 0 and dt < 1e-8 and dr < 1e-8 and st["n_edge"] == [0, 0]
        far = q["init"].copy()
        far[:3] += 500.0
        ok, x_ref = R.scan2map(c["map_corner"], c["map_surf"], q["corner"], q["surf"], far)
        rc, x, st = e.scan2map(q["corner"], q["surf"], far)
        assert rc == 0 and np.array_equal(x, x_ref) and np.array_equal(x, far)
    finally:
        e.close()


@pytest.mark.gpu
def test_cuda_scan2scan_equals_the_reference_matcher():
    from msf_loam_b200 import Engine
    from msf_loam_b200.engine import to_pcl
    P, (lc, lcr, ls, lsr, cs, cf), gt = _scan_pair()
    e = Engine()
    try:
        for init in (S.pose_identity(), gt):
            R.reset_logs()
            ok, x_ref = R.scan2scan(lc, lcr, ls, lsr, cs, cf, init)
            solves = R.solves()
            rc, x, st = e.scan2scan(to_pcl(lc, lcr), to_pcl(ls, lsr), cs, cf, init)
            assert ok is True and rc == 0
            dt, dr = S.pose_error(x, x_ref)
            assert dt < 1e-8 and dr < 1e-8, (dt, dr)
            assert st["n_edge"] == [s["n_edge"] for s in solves] and st["n_plane"] == [s["n_plane"] for s in solves]
        init = np.array([0.01, 0.02, 0.0, 0, 0, 0, 1.0])
        ok, x_ref = R.scan2scan(lc, lcr, ls, lsr, cs[:3], cf[:4], init)
        rc, x, st = e.scan2scan(to_pcl(lc, lcr), to_pcl(ls, lsr), cs[:3], cf[:4], init)
        assert ok is False and rc == 1 and np.array_equal(x, x_ref)
    finally:
        e.close()


@pytest.mark.gpu
def test_cuda_deskew_branch_equals_the_reference_matcher(vlp16_case):
    from msf_loam_b200 import Engine
    c = vlp16_case
    t, dq, dp = make_preintegration()
    e = Engine()
    try:
        e.set_submap(c["map_corner"], c["map_surf"])
        for q in c["queries"][:2]:
            ok, x_ref, _ = R.scan2map_deskew(c["map_corner"], c["map_surf"], q["corner"], q["surf"], t, dq, dp, VEL, GRAV, q["init"])
            rc, x, st = e.scan2map_deskew(q["corner"], q["surf"], t, dq, dp, VEL, GRAV, q["init"])
            assert ok is True and rc == 0
            dt, dr = S.pose_error(x, x_ref)
            # CUDA's acos / sin differ from glibc's in the last bit, which can flip the fp32 rounding of a kNN query
            assert dt < 1e-7 and dr < 1e-7, (dt, dr)
    finally:
        e.close()
