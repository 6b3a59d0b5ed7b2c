"""The REFERENCE's own factor code as the checker (rows a-8 / a-9 Plus / f-3 factors of SURVEY.md section 8).

oracle/_ref/libmsfl_ref.so contains lidar_factor.cc (:7-100) and pose_local_parameterization.cc (:6-27, through
Utility::deltaQ utility.h:8-31) compiled UNMODIFIED from the reference checkout (oracle/Makefile target `ref`; Eigen and
Ceres are absent from the image, so the sources see the stand-in headers of oracle/ref_stubs/ -- msfl_eigen_standin.h
says what that pins).  These tests check
  * the oracle's restatement of the four factors and of Plus against it (CPU),
  * the oracle's H / g / cost assembly against sums built from the reference's residuals and Jacobians (CPU),
  * the CUDA path (msfl_accumulate through the C ABI) against the same sums (GPU).
The prebuilt .so travels to the GPU box; /root/reference is only needed to (re)build it.
"""
import ctypes as C

import numpy as np
import pytest

import oracle as O
from msf_loam_b200 import synth as S

R = O.ref_lib()
pytestmark = pytest.mark.skipif(R is None, reason="no reference checkout and no prebuilt oracle/_ref library")

_D = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(_D)


def _arr(*vs):
    return [np.ascontiguousarray(v, dtype=np.float64) for v in vs]


def ref_factor(kind, pose, p, c, n):
    nres = 3 if kind == 0 else 1
    r, J = np.zeros(nres), np.zeros(nres * 7)
    a = _arr(pose, p, c, n)
    (R.msflref_edge_factor if kind == 0 else R.msflref_plane_factor)(*[_p(v) for v in a], _p(r), _p(J))
    return r, J.reshape(nres, 7)


def ref_factor_deskew(kind, pose, V, p, c, n, dp, dq, dt, G):
    nres = 3 if kind == 0 else 1
    r, J, Jb = np.zeros(nres), np.zeros(nres * 7), np.zeros(nres * 9)
    sb = np.zeros(9)
    sb[:3] = V
    a = _arr(pose, sb, p, c, n, dp, dq)
    g = _arr(G)[0]
    (R.msflref_edge_factor_deskew if kind == 0 else R.msflref_plane_factor_deskew)(
        *[_p(v) for v in a], C.c_double(dt), _p(g), _p(r), _p(J), _p(Jb))
    return r, J.reshape(nres, 7), Jb.reshape(nres, 9)


def ref_pose_plus(x, d):
    out = np.zeros(7)
    a = _arr(x, d)
    R.msflref_pose_plus(_p(a[0]), _p(a[1]), _p(out))
    return out


def _random_pose(rng, rot=0.6):
    return np.concatenate([rng.normal(scale=3, size=3), S.rotvec_to_quat(rng.normal(scale=rot, size=3))])


def test_reference_sizes_and_plus_jacobian():
    # GlobalSize 7 / LocalSize 6, ComputeJacobian = [I6; 0] (pose_local_parameterization.h:8-9, .cc:23-27): the local
    # Jacobian Ceres multiplies out is the first six columns of the factor's 7-column block
    assert R.msflref_pose_sizes() == 706
    J = np.zeros(42)
    R.msflref_pose_plus_jacobian(_p(np.zeros(7)), _p(J))
    assert np.array_equal(J.reshape(7, 6), np.eye(7, 6))


def test_oracle_factors_equal_the_reference_factors():
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(2000):
        pose = _random_pose(rng)
        p, c = rng.normal(scale=20, size=3), rng.normal(scale=20, size=3)
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        for kind, fn in ((0, O.edge_factor), (1, O.plane_factor)):
            r, J = fn(pose, p, c, n)
            rr, Jr = ref_factor(kind, pose, p, c, n)
            scale = 1.0 + np.abs(Jr).max()
            worst = max(worst, np.abs(r - rr).max() / scale, np.abs(J - Jr).max() / scale)
    assert worst <= 1e-14, worst  # measured: bit-equal


def test_oracle_deskew_factors_equal_the_reference_factors():
    rng = np.random.default_rng(6)
    L = O.lib()
    worst = 0.0
    for _ in range(2000):
        pose = _random_pose(rng)
        p, c = rng.normal(scale=20, size=3), rng.normal(scale=20, size=3)
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        dq, dp = S.rotvec_to_quat(rng.normal(scale=0.05, size=3)), rng.normal(scale=0.05, size=3)
        V, G, dt = rng.normal(scale=2, size=3), np.array([0.0, 0.0, 9.81]), float(rng.uniform(0, 0.1))
        for kind, fn in ((0, L.msflo_edge_factor_deskew), (1, L.msflo_plane_factor_deskew)):
            nres = 3 if kind == 0 else 1
            r, J = np.zeros(nres), np.zeros(nres * 7)
            a = _arr(pose, V, p, c, n, dp, dq)
            fn(*[_p(v) for v in a], C.c_double(dt), _p(G), _p(r), _p(J))
            rr, Jr, Jb = ref_factor_deskew(kind, pose, V, p, c, n, dp, dq, dt, G)
            scale = 1.0 + np.abs(Jr).max()
            worst = max(worst, np.abs(r - rr).max() / scale, np.abs(J.reshape(nres, 7) - Jr).max() / scale)
            # the speed-bias block is held constant by the reference (mapping_scan_matcher.cc:94): its Jacobian is
            # never used, but it is what the reference says it is: d r / d V = (N x or N^T) dt
            assert np.allclose(Jb[:, 3:], 0)
    assert worst <= 1e-14, worst


def test_oracle_pose_plus_equals_the_reference_plus():
    rng = np.random.default_rng(7)
    worst = 0.0
    for k in range(3000):
        x = _random_pose(rng)
        scale = [1.0, 1e-3, 1e-7, 1e-9][k % 4]  # both branches of deltaQ (theta < 1e-6: Taylor series)
        d = np.concatenate([rng.normal(size=3), rng.normal(scale=scale, size=3)])
        worst = max(worst, np.abs(O.pose_plus(x, d) - ref_pose_plus(x, d)).max())
    assert worst <= 1e-15 * 20, worst
    assert np.array_equal(ref_pose_plus(x, np.zeros(6))[:3], x[:3])


def _huber_sums_from_reference(corr, pose, a=0.1):
    """cost, H, g the way Ceres assembles them (corrector.cc with HuberLoss: rho'' <= 0 -> residual and Jacobian are
    scaled by sqrt(rho')), from the REFERENCE's residuals and Jacobians."""
    cost, H, g = 0.0, np.zeros((6, 6)), np.zeros(6)
    for c in corr:
        r, J = ref_factor(int(c[0]), pose, c[1:4], c[4:7], c[7:10])
        s = float(r @ r)
        if s > a * a:
            rho0, rho1 = 2 * a * np.sqrt(s) - a * a, a / np.sqrt(s)
        else:
            rho0, rho1 = s, 1.0
        cost += 0.5 * rho0
        Jl = J[:, :6] * np.sqrt(rho1)
        g += Jl.T @ (r * np.sqrt(rho1))
        H += Jl.T @ Jl
    return cost, H, g


def test_oracle_normal_equations_equal_the_reference_sums(vlp16_case):
    P = O.default_params()
    q = vlp16_case["queries"][1]
    corr, ne, npl, _ = O.associate_map(P, vlp16_case["map_corner"], vlp16_case["map_surf"], q["corner"], q["surf"], q["init"])
    assert ne > 100 and npl > 1000
    cost, H, g = O.accumulate(P, corr, q["init"])
    cost_r, H_r, g_r = _huber_sums_from_reference(corr, q["init"])
    assert abs(cost - cost_r) <= 1e-12 * cost_r
    assert np.abs(H - H_r).max() <= 1e-12 * np.abs(H_r).max()
    assert np.abs(g - g_r).max() <= 1e-12 * np.abs(g_r).max()


@pytest.mark.gpu
def test_cuda_normal_equations_equal_the_reference_sums(vlp16_case):
    from msf_loam_b200 import Engine
    P = O.default_params()
    eng = Engine()
    try:
        for qi in (0, 1):
            q = vlp16_case["queries"][qi]
            corr, ne, npl, _ = O.associate_map(P, vlp16_case["map_corner"], vlp16_case["map_surf"], q["corner"], q["surf"],
                                               q["init"])
            cost_g, H_g, g_g = eng.accumulate(corr[:, 1:4], corr[:, 4:10], ne, npl, q["init"])
            cost_r, H_r, g_r = _huber_sums_from_reference(corr, q["init"])
            assert abs(cost_g - cost_r) <= 1e-12 * cost_r
            assert np.abs(H_g - H_r).max() <= 1e-11 * np.abs(H_r).max()
            assert np.abs(g_g - g_r).max() <= 1e-11 * np.abs(g_r).max()
    finally:
        eng.close()


def test_stand_in_algebra_agrees_with_scipy():
    """The stand-in Eigen primitives under the compiled reference code (oracle/ref_stubs/msfl_eigen_standin.h) against an
    independent implementation: quaternion * vector and toRotationMatrix (through the factors' residuals / Jacobians),
    quaternion product + normalisation (through Plus), slerp (through GetDeltaQP) -- scipy.spatial.transform."""
    from scipy.spatial.transform import Rotation, Slerp
    from oracle import ref as RR
    rng = np.random.default_rng(8)
    for _ in range(200):
        pose = _random_pose(rng)
        Rm = Rotation.from_quat(pose[3:]).as_matrix()
        p, c = rng.normal(scale=10, size=3), rng.normal(scale=10, size=3)
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        r, J = ref_factor(1, pose, p, c, n)
        assert abs(r[0] - n @ (Rm @ p + pose[:3] - c)) < 1e-12
        skew = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
        assert np.abs(J[0, :3] - n).max() < 1e-15 and np.abs(J[0, 3:6] + n @ (Rm @ skew)).max() < 1e-12 and J[0, 6] == 0
        r3, J3 = ref_factor(0, pose, p, c, n)
        assert np.abs(r3 - np.cross(n, Rm @ p + pose[:3] - c)).max() < 1e-12
        # Plus: q <- (q * exp(dtheta / 2)) normalised, p <- p + dp
        d = np.concatenate([rng.normal(size=3), rng.normal(scale=0.3, size=3)])
        y = ref_pose_plus(pose, d)
        q_expect = (Rotation.from_quat(pose[3:]) * Rotation.from_rotvec(d[3:])).as_quat()
        q_expect *= np.sign(q_expect @ y[3:])
        assert np.abs(y[:3] - pose[:3] - d[:3]).max() < 1e-15 and np.abs(y[3:] - q_expect).max() < 1e-14
        # TransformPoint: float -> double -> R p + t -> float
        x = rng.normal(scale=30, size=(1, 3)).astype(np.float32)
        got = RR.transform_point(pose, x)[0]
        want = (Rm @ x[0].astype(np.float64) + pose[:3])
        assert np.abs(got - want).max() <= 4e-6  # one float ulp at 60 m
    # slerp + lerp of GetDeltaQP
    t = np.linspace(0.0, 0.1, 11)
    rots = Rotation.from_rotvec(np.outer(t, [0.3, -0.2, 1.1]))
    dq, dp = rots.as_quat(), np.outer(t, [0.5, 0.1, -0.2])
    sl = Slerp(t, rots)
    for dt in rng.uniform(0.0, 0.0999, size=100):
        q, pp = RR.get_delta_qp(t, dq, dp, float(dt))
        qe = sl([dt]).as_quat()[0]
        qe *= np.sign(qe @ q)
        assert np.abs(q - qe).max() < 1e-14 and np.abs(pp - np.array([0.5, 0.1, -0.2]) * dt).max() < 1e-15
