"""Validation of the CPU oracle itself (no GPU).  The reference's own tests pin nothing on this path
(SURVEY.md 4), so the oracle is cross-checked against independent implementations: brute force and
scipy cKDTree for the k-NN, numpy eigh / lstsq for the fits, finite differences through Plus for
the analytic Jacobians, an independent numpy Levenberg-Marquardt for the solve, known-transform
recovery, and a numpy restatement of the curvature / VoxelGrid arithmetic."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

import oracle as O
from msf_loam_b200 import synth as S


def rand_pose(rng, t=1.0, r=0.5):
    return np.concatenate([rng.normal(scale=t, size=3), S.rotvec_to_quat(rng.normal(scale=r, size=3))])


# ---------------------------------------------------------------------------------- k-NN
def test_knn_matches_brute_force_and_ckdtree():
    rng = np.random.default_rng(0)
    pts = np.zeros((5000, 4), np.float32)
    pts[:, :3] = rng.uniform(-20, 20, size=(5000, 3))
    pts[100:200, :3] = pts[0:100, :3]  # exact duplicates -> distance ties, lowest index must win
    q = rng.uniform(-21, 21, size=(700, 3)).astype(np.float32)
    q[:50] = pts[:50, :3]
    for k in (1, 5):
        idx, d2 = O.knn(pts, q, k)
        idx_b, d2_b = O.knn(pts, q, k, brute=True)
        assert np.array_equal(idx, idx_b) and np.array_equal(d2, d2_b)
        assert np.all(np.diff(d2, axis=1) >= 0)
        dd, ii = cKDTree(pts[:, :3].astype(np.float64)).query(q.astype(np.float64), k=k)
        dd = dd.reshape(len(q), k)
        assert np.allclose(np.sqrt(d2), dd, rtol=1e-5, atol=1e-5)


def test_map_association_knn_matches_real_flann_kdtree_single(vlp16_case):
    """Pins the 5-NN of rows a-6/a-7 against an actual FLANN build: OpenCV bundles FLANN and exposes
    KDTreeSingleIndex (algorithm 4, the index pcl::KdTreeFLANN uses; leaf size 15 as in PCL) with the fp32
    ((dx^2 + dy^2) + dz^2) distance.  Same neighbours in the same order, and the same d5^2 < 1 gate
    (mapping_scan_matcher.cc:125-128 / :195-198), on the real VLP-16 case."""
    cv2 = pytest.importorskip("cv2")
    P = O.default_params()
    q = vlp16_case["queries"][0]
    _, _, _, kidx = O.associate_map(P, vlp16_case["map_corner"], vlp16_case["map_surf"], q["corner"], q["surf"], q["init"])
    n_c = q["corner"].shape[0]
    checked = 0
    for cloud_map, scan, rows in ((vlp16_case["map_corner"], q["corner"], kidx[:n_c]), (vlp16_case["map_surf"], q["surf"], kidx[n_c:])):
        m = np.ascontiguousarray(cloud_map[:, :3])
        x = np.ascontiguousarray(S.transform_cloud(q["init"], scan)[:, :3])  # TransformPoint: fp64 math, fp32 store
        index = cv2.flann_Index(m, dict(algorithm=4, leaf_max_size=15))
        idx, d2 = index.knnSearch(x, 5, params=dict(checks=-1, eps=0.0, sorted=True))
        gate = d2[:, 4] < np.float32(1.0)
        assert np.array_equal(gate, rows[:, 0] >= 0)
        # equal distances may come back in either order from a kd-tree: compare as sets there, exactly elsewhere
        tie = (np.diff(d2, axis=1) == 0).any(axis=1)
        ok = gate & ~tie
        assert np.array_equal(rows[ok], idx[ok])
        for i in np.nonzero(gate & tie)[0]:
            assert sorted(rows[i]) == sorted(idx[i])
        checked += int(ok.sum())
    assert checked > 3000


def test_map_association_knn_matches_real_flann_on_hdl64():
    """The same FLANN pin on the HDL-64E-shape case (BASELINE config 3: 80 x 60 m scene, ~13 k queries, ~43 k map points)."""
    cv2 = pytest.importorskip("cv2")
    from conftest import make_map_case
    case = make_map_case("hdl64", "room80", 5, 200)
    P = O.default_params()
    q = case["queries"][0]
    _, _, _, kidx = O.associate_map(P, case["map_corner"], case["map_surf"], q["corner"], q["surf"], q["init"])
    n_c = q["corner"].shape[0]
    checked = 0
    for cloud_map, scan, rows in ((case["map_corner"], q["corner"], kidx[:n_c]), (case["map_surf"], q["surf"], kidx[n_c:])):
        m = np.ascontiguousarray(cloud_map[:, :3])
        x = np.ascontiguousarray(S.transform_cloud(q["init"], scan)[:, :3])
        idx, d2 = cv2.flann_Index(m, dict(algorithm=4, leaf_max_size=15)).knnSearch(x, 5, params=dict(checks=-1, eps=0.0, sorted=True))
        gate = d2[:, 4] < np.float32(1.0)
        assert np.array_equal(gate, rows[:, 0] >= 0)
        ok = gate & ~(np.diff(d2, axis=1) == 0).any(axis=1)
        assert np.array_equal(rows[ok], idx[ok])
        checked += int(ok.sum())
    assert checked > 9000


def test_odometry_nearest_neighbour_matches_real_flann():
    """Row a-5: the 1-NN of every sharp / flat point in the last scan's less-sharp / less-flat cloud
    (odometry_scan_matcher.cc:84, :169; kd-trees built at :57-61) against FLANN's KDTreeSingleIndex, with the
    d^2 < 25 gate of :87 / :172.  The oracle reports the association [closest, second] / [closest, j, l]; its first
    entry is that 1-NN."""
    cv2 = pytest.importorskip("cv2")
    P = O.default_params()
    sc, traj = S.make_scene(), S.trajectory(2, seed=1)
    f = [O.extract_features(P, *S.raycast_scan(sc, "vlp16", traj[k], seed=1 + k), None) for k in range(2)]
    lc, ls = f[0]["full"][f[0]["idx_less_sharp"]], f[0]["full"][f[0]["idx_less_flat"]]
    cs, cf = f[1]["full"][f[1]["idx_sharp"]], f[1]["full"][f[1]["idx_flat"]]
    init = S.pose_identity()
    rc, x, logs, counts, assoc = O.scan2scan(P, lc, f[0]["ring"][f[0]["idx_less_sharp"]], ls, f[0]["ring"][f[0]["idx_less_flat"]],
                                            cs, cf, init)
    # `assoc` is the association of the LAST outer iteration; redo iteration 0 (identity guess) through the k-NN entry point
    first_sharp, _ = O.knn(lc, cs[:, :3], 1)
    first_flat, _ = O.knn(ls, cf[:, :3], 1)
    checked = 0
    for target, queries, mine in ((lc, cs, first_sharp), (ls, cf, first_flat)):
        index = cv2.flann_Index(np.ascontiguousarray(target[:, :3]), dict(algorithm=4, leaf_max_size=15))
        idx, d2 = index.knnSearch(np.ascontiguousarray(queries[:, :3]), 1, params=dict(checks=-1, eps=0.0, sorted=True))
        assert np.array_equal(idx[:, 0], mine[:, 0])
        assert (d2[:, 0] < np.float32(25.0)).all()  # every query of this pair passes the kDistanceSqThreshold gate
        checked += len(idx)
    assert checked > 500 and rc == 0


def test_atan2_overload_of_the_registration_block():
    """msf_loam_node.cc:131,139 call atan2 unqualified on floats; with <math.h> in the include graph (ROS / tf pull it in)
    libstdc++ resolves that to float atan2(float, float), with <cmath> alone to the promoted double version.  The oracle
    and k_feat_angles follow the float overload."""
    import os
    import shutil
    import subprocess
    import tempfile
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_harness", "atan2_overload.cc")
    with tempfile.TemporaryDirectory() as d:
        sizes = []
        for flag in ([], ["-DWITH_MATH_H"]):
            exe = os.path.join(d, "a" + str(len(flag)))
            subprocess.run([gxx, "-std=c++14", *flag, src, "-o", exe], check=True)
            sizes.append(int(subprocess.run([exe], capture_output=True, text=True, check=True).stdout))
    assert sizes == [8, 4]
    # and the oracle's relative time is the float-azimuth one: scan time of a point = float(-atan2f) based
    P = O.default_params()
    xyzi, ring = S.raycast_scan(S.make_scene(), "vlp16", S.trajectory(1)[0], seed=3)
    f = O.extract_features(P, xyzi, ring, None)
    j = 12345
    src_idx = np.nonzero((xyzi[:, :3] == f["full"][j, :3]).all(1))[0][0]
    first = xyzi[0]
    start = np.float64(-np.arctan2(np.float32(first[1]), np.float32(first[0]), dtype=np.float32))
    ori = np.float64(-np.arctan2(xyzi[src_idx, 1], xyzi[src_idx, 0], dtype=np.float32))
    rel = np.fmod(ori - start + 2 * np.pi, 2 * np.pi)
    t = np.float32(rel / (2 * np.pi) * 0.1)
    assert abs(float(t) - float(f["full"][j, 3])) <= 1.6e-8 or abs(float(t) + 0.1 - float(f["full"][j, 3])) <= 1.6e-8


def test_knn_fewer_points_than_k():
    pts = np.zeros((3, 4), np.float32)
    pts[:, 0] = [0, 1, 2]
    idx, d2 = O.knn(pts, np.zeros((1, 3), np.float32), 5)
    assert list(idx[0]) == [0, 1, 2, -1, -1] and np.isinf(d2[0, 3])


# ---------------------------------------------------------------------------------- dense kernels
def test_sym_eig3_matches_eigh():
    rng = np.random.default_rng(1)
    for _ in range(200):
        A = rng.normal(size=(5, 3)) * rng.uniform(0.01, 3, size=3)
        C = A.T @ A
        ev, V = O.sym_eig3(C)
        w, U = np.linalg.eigh(C)
        assert np.allclose(ev, w, rtol=1e-10, atol=1e-12 * w.max())
        assert abs(abs(V[:, 2] @ U[:, 2]) - 1) < 1e-8
        assert np.allclose(C @ V, V * ev, atol=1e-9 * w.max())


def test_lstsq_5x3_matches_numpy():
    rng = np.random.default_rng(2)
    for _ in range(200):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        base = rng.normal(size=(5, 3)) * 0.5
        A = base - np.outer(base @ n, n) + n * rng.uniform(2, 40) + rng.normal(scale=0.01, size=(5, 3))
        b = -np.ones(5)
        x = O.lstsq_5x3(A, b)
        x_np = np.linalg.lstsq(A, b, rcond=None)[0]
        assert np.allclose(x, x_np, rtol=1e-8, atol=1e-10)


# ---------------------------------------------------------------------------------- factors
def _fd_jacobian(fun, pose, eps=1e-6):
    r0 = fun(pose)
    J = np.zeros((len(r0), 6))
    for k in range(6):
        d = np.zeros(6); d[k] = eps
        J[:, k] = (fun(O.pose_plus(pose, d)) - fun(O.pose_plus(pose, -d))) / (2 * eps)
    return J


def test_factor_jacobians_match_finite_differences_through_plus():
    rng = np.random.default_rng(3)
    for _ in range(20):
        pose = rand_pose(rng)
        p, a = rng.normal(scale=5, size=3), rng.normal(scale=5, size=3)
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        r, J = O.edge_factor(pose, p, a, n)
        Jfd = _fd_jacobian(lambda x: O.edge_factor(x, p, a, n)[0], pose)
        assert np.abs(J[:, :6] - Jfd).max() < 1e-6 and np.all(J[:, 6] == 0)
        r, J = O.plane_factor(pose, p, a, n)
        Jfd = _fd_jacobian(lambda x: O.plane_factor(x, p, a, n)[0], pose)
        assert np.abs(J[:, :6] - Jfd).max() < 1e-6 and J[0, 6] == 0


def test_factor_residual_definitions():
    """r_edge = N x (Q p + P - C), r_plane = N . (Q p + P - C)  (lidar_factor.cc:12,32)."""
    rng = np.random.default_rng(4)
    pose = rand_pose(rng)
    R = S.quat_to_R(pose[3:])
    p, c = rng.normal(size=3), rng.normal(size=3)
    n = rng.normal(size=3); n /= np.linalg.norm(n)
    x = R @ p + pose[:3] - c
    assert np.allclose(O.edge_factor(pose, p, c, n)[0], np.cross(n, x), atol=1e-14)
    assert np.allclose(O.plane_factor(pose, p, c, n)[0], n @ x, atol=1e-14)


def test_pose_plus_is_right_multiplication():
    rng = np.random.default_rng(5)
    x = rand_pose(rng)
    d = rng.normal(scale=0.1, size=6)
    y = O.pose_plus(x, d)
    q = S.quat_mul(x[3:], S.rotvec_to_quat(d[3:]))
    assert np.allclose(y[:3], x[:3] + d[:3]) and np.allclose(y[3:], q / np.linalg.norm(q), atol=1e-15)
    tiny = np.array([0, 0, 0, 1e-9, -2e-9, 5e-10])  # Taylor branch of deltaQ (utility.h:19-24)
    y = O.pose_plus(x, tiny)
    q = S.quat_mul(x[3:], np.concatenate([0.5 * tiny[3:], [1.0]]))
    assert np.allclose(y[3:], q / np.linalg.norm(q), atol=1e-15)


# ---------------------------------------------------------------------------------- LM
def _synthetic_corr(rng, pose_true, n_edge=150, n_plane=600, noise=0.0, outliers=0):
    R = S.quat_to_R(pose_true[3:])
    corr = []
    for k in range(n_edge + n_plane):
        p = rng.uniform(-15, 15, size=3)
        x = R @ p + pose_true[:3]
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        if k < n_edge:
            a = x + n * rng.uniform(-1, 1) + rng.normal(scale=noise, size=3)
        else:
            t = rng.normal(size=3); t -= (t @ n) * n
            a = x + t + n * rng.normal(scale=noise)
        if k % 40 == 0 and outliers:
            a = a + rng.normal(scale=2.0, size=3)
        corr.append(np.concatenate([[0.0 if k < n_edge else 1.0], p, a, n]))
    return np.array(corr)


def _numpy_lm(P, corr, pose):
    """Independent numpy restatement of the same Ceres semantics via explicit residual/Jacobian stacks."""
    def evaluate(x):
        rs, Js, cost = [], [], 0.0
        for c in corr:
            r, J = (O.edge_factor if c[0] == 0 else O.plane_factor)(x, c[1:4], c[4:7], c[7:10])
            s = float(r @ r)
            if s > P.huber_a ** 2:
                rho0, rho1 = 2 * P.huber_a * np.sqrt(s) - P.huber_a ** 2, P.huber_a / np.sqrt(s)
            else:
                rho0, rho1 = s, 1.0
            cost += 0.5 * rho0
            rs.append(np.sqrt(rho1) * r); Js.append(np.sqrt(rho1) * J[:, :6])
        return cost, np.concatenate(rs), np.vstack(Js)
    x = np.array(pose, dtype=np.float64)
    cost, r, J = evaluate(x)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0)))
    radius, nu, reuse, diag = P.initial_radius, 2.0, False, None
    for _ in range(P.max_num_iterations):
        Js = J * scale
        H, g = Js.T @ Js, Js.T @ r
        if not reuse:
            diag = np.clip(np.diag(H), P.min_lm_diagonal, P.max_lm_diagonal)
        y = np.linalg.solve(H + np.diag(diag / radius), -g)
        model = -(y @ g + 0.5 * y @ H @ y)
        xc = O.pose_plus(x, y * scale)
        cost_c = evaluate(xc)[0]
        if np.linalg.norm(x - xc) <= P.parameter_tolerance * (np.linalg.norm(x) + P.parameter_tolerance):
            break
        if abs(cost - cost_c) <= P.function_tolerance * cost:
            break
        rho = (cost - cost_c) / model
        if rho > P.min_relative_decrease:
            x = xc
            cost, r, J = evaluate(x)
            radius = min(P.max_radius, radius / max(1 / 3, 1 - (2 * rho - 1) ** 3)); nu = 2.0; reuse = False
        else:
            radius /= nu; nu *= 2; reuse = True
    return x


def test_lm_agrees_with_independent_numpy_lm():
    rng = np.random.default_rng(6)
    P = O.default_params()
    truth = rand_pose(rng, 2.0, 0.3)
    corr = _synthetic_corr(rng, truth, n_edge=60, n_plane=200, noise=0.02, outliers=1)
    init = S.perturb_pose(truth, rng, 0.2, 2.0)
    x, log = O.lm_solve(P, corr, init)
    x_np = _numpy_lm(P, corr, init)
    assert np.abs(x - x_np).max() < 1e-10
    costs = [it["cost"] for it in log["iters"]]
    assert all(b <= a for a, b in zip(costs, costs[1:]))  # monotone
    assert log["n_attempts"] <= P.max_num_iterations


def test_lm_converged_minimiser_matches_scipy_least_squares_with_huber_loss():
    """A different solver on the same objective: scipy.optimize.least_squares (trust-region reflective, its own Huber
    loss, finite-difference Jacobian) over a minimal 6-vector around the oracle's answer.  Ceres applies the loss per
    residual BLOCK (s = |r|^2 of the 3-vector edge residual), so an edge block enters scipy as the single residual |r|.
    The converged poses must agree far below the 1e-4 parity bound."""
    from scipy.optimize import least_squares
    rng = np.random.default_rng(16)
    P = O.default_params(max_num_iterations=30, early_exit=0)
    truth = rand_pose(rng, 2.0, 0.3)
    corr = _synthetic_corr(rng, truth, n_edge=60, n_plane=200, noise=0.03, outliers=1)
    x, log = O.lm_solve(P, corr, S.perturb_pose(truth, rng, 0.2, 2.0))

    def residuals(d):
        xd = O.pose_plus(x, d)
        out = []
        for c in corr:
            r, _ = (O.edge_factor if c[0] == 0 else O.plane_factor)(xd, c[1:4], c[4:7], c[7:10])
            out.append(np.linalg.norm(r) if c[0] == 0 else r[0])
        return np.array(out)

    sol = least_squares(residuals, np.zeros(6), loss="huber", f_scale=P.huber_a, xtol=1e-15, ftol=1e-15, gtol=1e-15,
                        x_scale=1.0, diff_step=1e-7)
    assert sol.cost <= log["final_cost"] * (1 + 1e-9)           # scipy may polish, never by much
    assert abs(sol.cost - log["final_cost"]) <= 1e-7 * log["final_cost"]
    dt, dr = S.pose_error(O.pose_plus(x, sol.x), x)
    assert dt < 2e-6 and dr < 2e-6
    n_out = int((np.abs(residuals(np.zeros(6))) > P.huber_a).sum())
    assert n_out >= 3  # the loss is actually active at the optimum


def test_lm_recovers_known_transform_noise_free():
    rng = np.random.default_rng(7)
    P = O.default_params(max_num_iterations=15, early_exit=0)
    truth = rand_pose(rng, 2.0, 0.3)
    corr = _synthetic_corr(rng, truth)
    x, log = O.lm_solve(P, corr, S.perturb_pose(truth, rng, 0.10, 1.0))
    dt, dr = S.pose_error(x, truth)
    assert dt < 1e-6 and dr < 1e-6


def test_lm_accumulate_consistency_and_huber():
    rng = np.random.default_rng(8)
    P = O.default_params()
    truth = rand_pose(rng)
    corr = _synthetic_corr(rng, truth, 30, 80, noise=0.3)
    cost, H, g = O.accumulate(P, corr, truth)
    assert np.allclose(H, H.T) and np.all(np.linalg.eigvalsh(H) > -1e-9)
    # cost = 1/2 sum rho(|r|^2) with Huber(0.1)
    ref = 0.0
    for c in corr:
        r = (O.edge_factor if c[0] == 0 else O.plane_factor)(truth, c[1:4], c[4:7], c[7:10])[0]
        s = float(r @ r)
        ref += 0.5 * (s if s <= 0.01 else 0.2 * np.sqrt(s) - 0.01)
    assert abs(cost - ref) < 1e-12 * max(1.0, ref)


def test_lm_no_correspondences_is_a_noop():
    P = O.default_params()
    x0 = np.array([1.0, 2, 3, 0, 0, 0, 1])
    x, log = O.lm_solve(P, np.zeros((0, 10)), x0)
    assert np.array_equal(x, x0) and log["n_attempts"] == 0


# ---------------------------------------------------------------------------------- extraction / voxel grid
def test_curvature_and_feature_invariants():
    P = O.default_params()
    xyzi, ring = S.raycast_scan(S.make_scene(), "vlp16", S.trajectory(1)[0], seed=1)
    f = O.extract_features(P, xyzi, ring, None)
    full, curv, label = f["full"], f["curvature"], f["label"]
    assert np.all(np.diff(f["ring"].astype(int)) >= 0)          # ring-major
    # numpy restatement of the fp32 11-tap sum (msf_loam_node.cc:213-236)
    X = full[:, :3]
    n = len(X)
    acc = X[0:n - 10].copy()
    for k in range(1, 5):
        acc = acc + X[k:n - 10 + k]
    acc = acc - np.float32(10) * X[5:n - 5]
    for k in range(6, 11):
        acc = acc + X[k:n - 10 + k]
    ref = (acc.astype(np.float64) ** 2).sum(axis=1).astype(np.float32)
    assert np.array_equal(curv[5:n - 5], ref)
    # picks respect the 0.1 threshold and the per-sector caps
    assert np.all(curv[f["idx_sharp"]] > 0.1) and np.all(curv[f["idx_flat"]] < 0.1)
    n_rings = int(f["ring"].max()) + 1
    assert len(f["idx_sharp"]) <= n_rings * 6 * 2 and len(f["idx_less_sharp"]) <= n_rings * 6 * 20
    assert len(f["idx_flat"]) <= n_rings * 6 * 4
    assert set(f["idx_sharp"]).issubset(set(f["idx_less_sharp"]))
    # less-flat = FLAT/UNKNOWN at the time its sector was closed; the next sector may later re-label up to
    # 5 of its trailing points LESS_SHARP (the order dependence noted in SURVEY.md a-3)
    late = ~np.isin(label[f["idx_less_flat"]], [0, 3])
    assert late.sum() <= 5 * 6 * n_rings and late.mean() < 0.01
    # relative time in [0, scan_period) and increasing inside a ring (clockwise input)
    t = full[:, 3]
    assert t.min() >= 0 and t.max() < 0.1 + 1e-3
    for r in range(n_rings):
        tr = t[f["ring"] == r]
        assert np.all(np.diff(tr) > -1e-6)


def test_extract_does_not_depend_on_input_interleaving():
    P = O.default_params()
    xyzi, ring = S.raycast_scan(S.make_scene(), "vlp16", S.trajectory(1)[0], seed=2)
    f1 = O.extract_features(P, xyzi, ring, None)
    order = np.argsort(ring, kind="stable")
    order = np.concatenate([order[:1], order[1:]])
    # same first point (start_ori) requires the first fired point to stay first
    first = 0
    rest = order[order != first]
    perm = np.concatenate([[first], rest])
    f2 = O.extract_features(P, xyzi[perm], ring[perm], None)
    assert np.array_equal(f1["idx_flat"], f2["idx_flat"]) and np.array_equal(f1["full"], f2["full"])


def test_voxel_grid_matches_numpy_restatement():
    rng = np.random.default_rng(9)
    pts = rng.uniform(-10, 10, size=(20000, 4)).astype(np.float32)
    for leaf in (0.2, 0.4, 3.0):
        a = O.voxel_grid(pts, leaf)
        b = S.voxel_grid_np(pts, leaf)
        assert a.shape == b.shape and np.array_equal(a, b)
    one = O.voxel_grid(pts[:1], 0.2)
    assert np.array_equal(one, pts[:1])
    # centroid property
    a = O.voxel_grid(pts, 1.0)
    assert abs(a[:, :3].mean() - pts[:, :3].mean()) < 0.05


# ---------------------------------------------------------------------------------- end to end
def test_scan2map_localises_and_batch_matches_single(vlp16_case):
    P = O.default_params()
    c = vlp16_case
    poses = []
    for q in c["queries"]:
        x, logs, counts = O.scan2map(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
        dt, dr = S.pose_error(x, q["gt"])
        assert dt < 0.03 and dr < 0.005 and counts[0, 1] > 1000
        poses.append(x)
    co = np.concatenate([[0], np.cumsum([q["corner"].shape[0] for q in c["queries"]])])
    so = np.concatenate([[0], np.cumsum([q["surf"].shape[0] for q in c["queries"]])])
    xb = O.scan2map_batch(P, c["map_corner"], c["map_surf"], np.concatenate([q["corner"] for q in c["queries"]]), co,
                          np.concatenate([q["surf"] for q in c["queries"]]), so,
                          np.stack([q["init"] for q in c["queries"]]), n_threads=3)
    assert np.array_equal(xb, np.stack(poses))


def test_scan2scan_config1_cpu_plumbing():
    """BASELINE config 1: single VLP-16 scan pair, scan-to-scan odometry on the CPU path."""
    P = O.default_params()
    sc, traj = S.make_scene(), S.trajectory(2, seed=1)
    f = [O.extract_features(P, *S.raycast_scan(sc, "vlp16", traj[k], seed=1 + k), None) for k in range(2)]
    rc, x, logs, counts, assoc = O.scan2scan(
        P, f[0]["full"][f[0]["idx_less_sharp"]], f[0]["ring"][f[0]["idx_less_sharp"]],
        f[0]["full"][f[0]["idx_less_flat"]], f[0]["ring"][f[0]["idx_less_flat"]],
        f[1]["full"][f[1]["idx_sharp"]], f[1]["full"][f[1]["idx_flat"]], S.pose_identity())
    gt = S.pose_mul(S.pose_inv(traj[0]), traj[1])
    dt, dr = S.pose_error(x, gt)
    assert rc == 0 and dt < 0.05 and dr < 0.01 and counts.sum() > 500
