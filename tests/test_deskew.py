"""SURVEY.md 8f row 3 -- the IMU-initialised branch of MatchScan2Map (Deskew factors).
CPU: the oracle's deskew factors against finite differences and against the plain branch.
GPU: msfl_scan2map_deskew against the oracle."""
import numpy as np
import pytest

import oracle as O
from msf_loam_b200 import synth as S


def make_preintegration(rng=None, rate_hz=400.0, span=0.105, omega=(0.02, -0.03, 0.25), acc=(0.4, -0.2, 0.1), v0=(0.05, 0.0, 0.0)):
    """Synthetic IntegrationBase buffers: constant angular rate / acceleration over one scan period."""
    t = np.arange(0.0, span + 1e-9, 1.0 / rate_hz)
    dq = np.stack([S.rotvec_to_quat(np.array(omega) * ti) for ti in t])
    dp = np.stack([np.array(v0) * ti + 0.5 * np.array(acc) * ti * ti for ti in t])
    return t, dq, dp


VEL = (0.3, 0.1, -0.02)
GRAV = (0.0, 0.0, 9.81)


def test_deskew_factor_jacobians_and_reduction_to_plain():
    rng = np.random.default_rng(0)
    import ctypes as C
    L = O.lib()
    for _ in range(10):
        pose = np.concatenate([rng.normal(size=3), S.rotvec_to_quat(rng.normal(scale=0.4, size=3))])
        p, c = rng.normal(scale=5, size=3), rng.normal(scale=5, size=3)
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        dq = S.rotvec_to_quat(rng.normal(scale=0.05, size=3)); dp = rng.normal(scale=0.05, size=3)
        dt = float(rng.uniform(0, 0.1))
        V, G = np.array(VEL), np.array(GRAV)

        def call(fn, nres, x):
            r = np.zeros(nres); J = np.zeros(nres * 7)
            a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, V, p, c, n, dp, dq)]
            fn(*[v.ctypes.data_as(C.POINTER(C.c_double)) for v in a], C.c_double(dt),
               G.ctypes.data_as(C.POINTER(C.c_double)), r.ctypes.data_as(C.POINTER(C.c_double)),
               J.ctypes.data_as(C.POINTER(C.c_double)))
            return r, J.reshape(nres, 7)

        for fn, nres in ((L.msflo_edge_factor_deskew, 3), (L.msflo_plane_factor_deskew, 1)):
            r, J = call(fn, nres, pose)
            Jfd = np.zeros((nres, 6))
            for k in range(6):
                d = np.zeros(6); d[k] = 1e-6
                Jfd[:, k] = (call(fn, nres, O.pose_plus(pose, d))[0] - call(fn, nres, O.pose_plus(pose, -d))[0]) / 2e-6
            assert np.abs(J[:, :6] - Jfd).max() < 1e-6
            # definition: N x / . (Q (dq p + dp) + V dt - g dt^2/2 + P - C)   (lidar_factor.cc:53,81)
            x = S.quat_to_R(pose[3:]) @ (S.quat_to_R(dq) @ p + dp) + V * dt - 0.5 * G * dt * dt + pose[:3] - c
            ref = np.cross(n, x) if nres == 3 else np.array([n @ x])
            assert np.allclose(r, ref, atol=1e-12)


def test_deskew_with_identity_motion_equals_plain_branch(vlp16_case):
    P = O.default_params()
    c, q = vlp16_case, vlp16_case["queries"][0]
    t = np.array([0.0, 0.2])
    dq = np.tile([0, 0, 0, 1.0], (2, 1)); dp = np.zeros((2, 3))
    rc, x, logs, counts, _ = O.scan2map_deskew(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], t, dq, dp,
                                               (0, 0, 0), (0, 0, 0), q["init"])
    x_ref, _, counts_ref = O.scan2map(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
    assert rc == 0 and np.array_equal(counts, counts_ref)
    dt_, dr_ = S.pose_error(x, x_ref)
    assert dt_ < 1e-9 and dr_ < 1e-9


def test_deskew_time_outside_preintegration_window_is_an_error(vlp16_case):
    P = O.default_params()
    c, q = vlp16_case, vlp16_case["queries"][0]
    t, dq, dp = make_preintegration(span=0.02)  # scan times reach ~0.1 s
    rc, *_ = O.scan2map_deskew(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], t, dq, dp, VEL, GRAV, q["init"])
    assert rc == -1


@pytest.mark.gpu
def test_cuda_deskew_branch_matches_oracle(vlp16_case):
    from msf_loam_b200 import Engine, MappingScanMatcher, MsflError, TimestampedPointCloud
    P = O.default_params()
    c = vlp16_case
    t, dq, dp = make_preintegration()
    e = Engine()
    e.set_submap(c["map_corner"], c["map_surf"])
    for q in c["queries"]:
        rc_ref, x_ref, logs, counts, kidx = O.scan2map_deskew(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"],
                                                              t, dq, dp, VEL, GRAV, q["init"])
        rc, x, st = e.scan2map_deskew(q["corner"], q["surf"], t, dq, dp, VEL, GRAV, q["init"])
        assert rc == rc_ref == 0
        dt_, dr_ = S.pose_error(x, x_ref)
        assert dt_ <= 1e-4 and dr_ <= 1e-4 and dt_ < 1e-7 and dr_ < 1e-7
        # counts may differ by a query or two: the interpolation uses acos/sin whose last bit differs
        # between glibc and CUDA, which can flip the fp32 rounding of a kNN query
        assert abs(st["n_edge"][0] - counts[0, 0]) <= 2 and abs(st["n_plane"][0] - counts[0, 1]) <= 4
        assert [l["n_attempts"] for l in st["lm"]] == [l["n_attempts"] for l in logs]
        # the deskewed solve really differs from the plain one (the motion is not negligible)
        x_plain = e.scan2map(q["corner"], q["surf"], q["init"])[1]
        assert S.pose_error(x, x_plain)[0] > 1e-3
    # reference-shaped call
    m = MappingScanMatcher(e)
    q = c["queries"][0]
    ok, pose = m.MatchScan2Map(TimestampedPointCloud(cloud_corner_less_sharp=c["map_corner"], cloud_surf_less_flat=c["map_surf"]),
                               TimestampedPointCloud(cloud_corner_less_sharp=q["corner"], cloud_surf_less_flat=q["surf"]),
                               True, q["init"], preintegration=(t, dq, dp), gravity_vector=GRAV, velocity=VEL)
    assert ok is True
    with pytest.raises(MsflError):  # point times outside the preintegration window (CHECK, scan_undistortion.cc:26)
        e.scan2map_deskew(q["corner"], q["surf"], t[:5], dq[:5], dp[:5], VEL, GRAV, q["init"])
    e.close()


@pytest.mark.gpu
def test_cuda_deskew_batch_equals_single_calls(vlp16_case):
    """msfl_scan2map_deskew_batch = B msfl_scan2map_deskew calls, bit for bit: every scan has its own preintegration
    table (different lengths, rates and motions), velocity and gravity; a scan whose times leave its window fails the
    whole call and no pose is written."""
    from msf_loam_b200 import Engine, MsflError
    c = vlp16_case
    tabs = [make_preintegration() + (VEL, GRAV),
            make_preintegration(rate_hz=200.0, span=0.11, omega=(-0.05, 0.02, -0.3), acc=(-0.3, 0.5, 0.0), v0=(0.0, 0.04, 0.01))
            + ((-0.2, 0.25, 0.05), (0.0, 0.3, 9.78)),
            make_preintegration(rate_hz=1000.0, span=0.12, omega=(0.0, 0.0, 0.0), acc=(0.0, 0.0, 0.0), v0=(0.0, 0.0, 0.0))
            + ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0)),
            make_preintegration() + (VEL, GRAV)]
    qs = [c["queries"][i % 3] for i in range(4)]
    e = Engine()
    try:
        e.set_submap(c["map_corner"], c["map_surf"])
        single = [e.scan2map_deskew(q["corner"], q["surf"], *t, q["init"]) for q, t in zip(qs, tabs)]
        rc, poses, st = e.scan2map_deskew_batch([q["corner"] for q in qs], [q["surf"] for q in qs], tabs,
                                                np.stack([q["init"] for q in qs]))
        assert rc == 0
        for b, (rc1, x1, st1) in enumerate(single):
            assert rc1 == 0 and np.array_equal(poses[b], x1), b
            assert st[b]["n_edge"] == st1["n_edge"] and st[b]["n_plane"] == st1["n_plane"]
            assert [l["n_attempts"] for l in st[b]["lm"]] == [l["n_attempts"] for l in st1["lm"]]
        # scans 0 and 3 are the same problem at different batch positions
        assert np.array_equal(poses[0], poses[3])
        # identity motion, zero velocity / gravity = the plain branch (scan 2)
        x_plain = e.scan2map(qs[2]["corner"], qs[2]["surf"], qs[2]["init"])[1]
        assert S.pose_error(poses[2], x_plain)[0] < 1e-9
        # against the oracle
        P = O.default_params()
        for b in (1, 3):
            _, x_ref, _, _, _ = O.scan2map_deskew(P, c["map_corner"], c["map_surf"], qs[b]["corner"], qs[b]["surf"], *tabs[b],
                                                  qs[b]["init"])
            dt_, dr_ = S.pose_error(poses[b], x_ref)
            assert dt_ < 1e-7 and dr_ < 1e-7
        bad = list(tabs)
        bad[1] = (tabs[1][0][:5], tabs[1][1][:5], tabs[1][2][:5]) + tabs[1][3:]
        init = np.stack([q["init"] for q in qs])
        with pytest.raises(MsflError, match="scan 1"):
            e.scan2map_deskew_batch([q["corner"] for q in qs], [q["surf"] for q in qs], bad, init)
    finally:
        e.close()
