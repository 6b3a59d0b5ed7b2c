"""Parity of the CUDA feature extraction / VoxelGrid / scan-to-scan path against the CPU oracle.

Bit-exact: ring-major order, rings, curvature (fp32), labels, the four index lists, voxel-grid
centroids, odometry associations.  Relative time (stored in `intensity`): the reference's azimuth is a
FLOAT atan2 (msf_loam_node.cc:131,139 resolve to the float overload, see oracle/ref_harness/atan2_overload.cc);
libm's atan2f is only defined to 1 ulp (2.4e-7 rad = 3.8e-9 s of scan time), and the result is rounded to float
once more -> compared to 1.6e-8 s (two float ulps at the end of the scan).  Poses: <= 1e-4 m / 1e-4 rad (north_star).
"""
import numpy as np
import pytest

import oracle as O
from msf_loam_b200 import (Engine, OdometryScanMatcher, ScanRegistration, TimestampedPointCloud,
                           MsflError, to_pcl)
from msf_loam_b200 import synth as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = Engine()
    yield e
    e.close()


def _check_features(f, g, check_time=True):
    assert f["full"].shape == g["full"].shape
    assert np.array_equal(f["ring"], g["ring"])
    assert np.array_equal(f["full"][:, :3], g["full"][:, :3])
    if check_time:
        assert np.abs(f["full"][:, 3] - g["full"][:, 3]).max() <= 1.6e-8
    assert np.array_equal(f["curvature"], g["curvature"])
    assert np.array_equal(f["label"], g["label"])
    for k in ("idx_sharp", "idx_less_sharp", "idx_flat", "idx_less_flat"):
        assert np.array_equal(f[k], g[k]), k


@pytest.mark.parametrize("sensor,scene", [("vlp16", "room40"), ("hdl64", "room80")])
def test_extract_features_bit_exact(eng, sensor, scene):
    P = O.default_params()
    sc = S.make_scene(scene)
    traj = S.trajectory(3)
    T = np.array([0.1, -0.2, 0.3, 0.0, 0.0, np.sin(0.05), np.cos(0.05)])
    for k in range(2):
        xyzi, ring = S.raycast_scan(sc, sensor, traj[k], seed=40 + k)
        f = O.extract_features(P, xyzi, ring, T)
        g = eng.extract_features(xyzi, ring, T)
        assert len(f["idx_sharp"]) > 50 and len(f["idx_flat"]) > 100
        _check_features(f, g)


def test_extract_features_os1_128_all_rings_bit_exact(eng):
    """128 rings = kMaxScanNum (msf_loam_node.cc:79): every CTA of the pick kernel is busy, ring 127 included."""
    P = O.default_params()
    xyzi, ring = S.raycast_scan(S.make_scene("room80"), "os1-128", S.trajectory(1)[0], seed=77)
    assert ring.max() == 127
    f = O.extract_features(P, xyzi, ring, None)
    g = eng.extract_features(xyzi, ring, None)
    _check_features(f, g)


@pytest.mark.parametrize("n_ring0", [5000, 13000, 30000])
def test_extract_long_rings_take_the_unstaged_and_batched_sort_paths(eng, n_ring0):
    """The pick kernel stages rings of <= 4096 points in shared memory and sorts all sectors of a ring at once when
    they fit 8192 keys: 5000 points -> global-memory flags, one sort batch; 13000 -> sectors of 2165 points sorted in
    batches of 4 + 2; 30000 -> sectors of 4998 points, one sector per batch.  A second, ordinary ring rides along."""
    rng = np.random.default_rng(n_ring0)
    def ring_cloud(n, rad, z):
        # clockwise azimuth (README.md:56-58), noisy radius with corners (a rounded square) so both feature kinds exist
        az = -np.linspace(0.0, 2 * np.pi, n, endpoint=False)
        sigma = np.where((np.floor(-az * 16 / (2 * np.pi)).astype(int) % 2) == 0, 0.04, 0.002)  # rough and smooth stretches
        rr = rad / np.maximum(np.abs(np.cos(az)), np.abs(np.sin(az))) + rng.normal(0, 1.0, n) * sigma
        return np.stack([rr * np.cos(az), rr * np.sin(az), np.full(n, z), np.zeros(n)], axis=1).astype(np.float32)
    xyzi = np.concatenate([ring_cloud(n_ring0, 8.0, -0.5), ring_cloud(1800, 6.0, 0.4)])
    ring = np.concatenate([np.zeros(n_ring0, np.uint16), np.ones(1800, np.uint16)])
    P = O.default_params()
    f = O.extract_features(P, xyzi, ring, None)
    g = eng.extract_features(xyzi, ring, None)
    assert len(f["idx_less_sharp"]) > 100 and len(f["idx_flat"]) > 24
    _check_features(f, g)


@pytest.mark.parametrize("shape", [0, 1, 2])
def test_extract_pick_shapes_agree(shape, monkeypatch):
    """k_feat_pick runs in one of three shapes (512 threads / 2048 staged ring points / 4096 sort keys ... 1024 / 4096 /
    8192), chosen from the rings of the engine's previous extraction; a sector that does not fit the keys re-runs the
    largest shape.  Every shape gives the oracle's lists: a VLP-16 scan, a batch, and a 30 000-point ring (sectors of
    4 998 points: overflow of the two smaller shapes -> retry)."""
    monkeypatch.setenv("MSFL_PICK_SHAPE", str(shape))
    e = Engine()
    P = O.default_params()
    sc = S.make_scene()
    traj = S.trajectory(2)
    scans = [S.raycast_scan(sc, "vlp16", traj[k], seed=300 + k) for k in range(2)]
    for xyzi, ring in scans:
        _check_features(O.extract_features(P, xyzi, ring, None), e.extract_features(xyzi, ring, None))
    fb = e.extract_features_batch([x for x, _ in scans], [r for _, r in scans], None)
    for (xyzi, ring), g in zip(scans, fb):
        _check_features(O.extract_features(P, xyzi, ring, None), g)
    rng = np.random.default_rng(5)
    n = 30000
    az = -np.linspace(0.0, 2 * np.pi, n, endpoint=False)
    rr = 8.0 / np.maximum(np.abs(np.cos(az)), np.abs(np.sin(az))) + rng.normal(0, 0.02, n)
    xyzi = np.stack([rr * np.cos(az), rr * np.sin(az), np.full(n, -0.5), np.zeros(n)], axis=1).astype(np.float32)
    ring = np.zeros(n, np.uint16)
    _check_features(O.extract_features(P, xyzi, ring, None), e.extract_features(xyzi, ring, None))


def test_extract_ring_major_input_invalid_points_and_ragged_rings(eng):
    """ring-major input, NaN / too-close points removed, rings with < 12 points skipped."""
    P = O.default_params()
    sc = S.make_scene()
    xyzi, ring = S.raycast_scan(sc, "vlp16", S.trajectory(1)[0], seed=7)
    order = np.argsort(ring, kind="stable")
    xyzi, ring = xyzi[order].copy(), ring[order].copy()
    rng = np.random.default_rng(3)
    bad = rng.choice(len(xyzi), 300, replace=False)
    xyzi[bad[:100], 0] = np.nan
    xyzi[bad[100:200], :3] *= 1e-3          # inside min_range
    xyzi[bad[200:], 2] = np.inf
    keep = ~((ring == 4) & (np.arange(len(ring)) % 200 != 0))  # ring 4 keeps ~9 points -> skipped (:252)
    xyzi, ring = xyzi[keep], ring[keep]
    f = O.extract_features(P, xyzi, ring, None)
    g = eng.extract_features(xyzi, ring, None)
    fin = np.isfinite(xyzi[:, :3]).all(axis=1)
    with np.errstate(invalid="ignore"):
        close = np.sqrt((xyzi[:, :3].astype(np.float32) ** 2).sum(axis=1)) < 0.3
    assert f["full"].shape[0] == int((fin & ~close).sum()) < len(xyzi) - 200
    _check_features(f, g)
    assert not np.any(f["ring"][f["idx_flat"]] == 4)


def test_extract_errors(eng):
    xyzi = np.ones((50, 4), np.float32)
    with pytest.raises(MsflError):  # ring >= 128 (CHECK_LT, msf_loam_node.cc:136)
        eng.extract_features(xyzi, np.full(50, 200, np.uint16))
    with pytest.raises(MsflError):  # no valid point (CHECK_GT(_N, 0), :200)
        eng.extract_features(np.zeros((50, 4), np.float32), np.zeros(50, np.uint16))


def test_voxel_grid_bit_exact(eng):
    P = O.default_params()
    sc = S.make_scene()
    xyzi, ring = S.raycast_scan(sc, "vlp16", S.trajectory(1)[0], seed=9)
    f = O.extract_features(P, xyzi, ring, None)
    for cloud, leaf in ((f["full"][f["idx_less_flat"]], 0.4), (f["full"][f["idx_less_sharp"]], 0.2),
                        (f["full"], 0.4), (f["full"][:1], 0.2)):
        a = O.voxel_grid(cloud, leaf)
        b = eng.voxel_grid(cloud, leaf)
        assert a.shape == b.shape and np.array_equal(a, b)
        assert np.array_equal(a, S.voxel_grid_np(cloud, leaf))
    assert eng.voxel_grid(np.zeros((0, 4), np.float32), 0.4).shape == (0, 4)


def _scan_pair(seed0=100):
    P = O.default_params()
    sc = S.make_scene()
    traj = S.trajectory(3)
    out = []
    for k in range(2):
        xyzi, ring = S.raycast_scan(sc, "vlp16", traj[k], seed=seed0 + k)
        out.append(O.extract_features(P, xyzi, ring, None))
    gt = S.pose_mul(S.pose_inv(traj[0]), traj[1])
    return P, out[0], out[1], gt


def test_scan2scan_association_bit_exact(eng):
    P, f0, f1, gt = _scan_pair()
    lc, lcr = f0["full"][f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]]
    ls, lsr = f0["full"][f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]]
    cs, cf = f1["full"][f1["idx_sharp"]], f1["full"][f1["idx_flat"]]
    for init in (S.pose_identity(), gt):
        rc, x, logs, counts, assoc = O.scan2scan(P, lc, lcr, ls, lsr, cs, cf, init)
        got = eng.associate_scan(to_pcl(lc, lcr), to_pcl(ls, lsr), cs, cf, init)
        assert np.array_equal(got, assoc)
        assert (assoc >= 0).sum() > 800


def test_scan2scan_pose_parity_config1(eng):
    """BASELINE config 1: single VLP-16 scan pair, scan-to-scan odometry."""
    P, f0, f1, gt = _scan_pair()
    lc, lcr = f0["full"][f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]]
    ls, lsr = f0["full"][f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]]
    cs, cf = f1["full"][f1["idx_sharp"]], f1["full"][f1["idx_flat"]]
    rc_ref, x_ref, logs, counts, _ = O.scan2scan(P, lc, lcr, ls, lsr, cs, cf, S.pose_identity())
    rc, x, st = eng.scan2scan(to_pcl(lc, lcr), to_pcl(ls, lsr), cs, cf, S.pose_identity())
    assert rc == rc_ref == 0
    dt, dr = S.pose_error(x, x_ref)
    assert dt <= 1e-4 and dr <= 1e-4
    assert dt < 1e-8 and dr < 1e-8
    assert st["n_edge"] == list(counts[:, 0]) and st["n_plane"] == list(counts[:, 1])
    for lg, lr in zip(st["lm"], logs):
        assert lg["n_attempts"] == lr["n_attempts"] and lg["termination"] == lr["termination"]
    dt_gt, dr_gt = S.pose_error(x, gt)
    assert dt_gt < 0.05 and dr_gt < 0.01


def test_scan2scan_too_few_correspondences(eng):
    """< 10 correspondences -> false, pose untouched (odometry_scan_matcher.cc:262-267)."""
    P, f0, f1, gt = _scan_pair()
    lc, lcr = f0["full"][f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]]
    ls, lsr = f0["full"][f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]]
    cs, cf = f1["full"][f1["idx_sharp"]][:3], f1["full"][f1["idx_flat"]][:4]
    init = np.array([0.01, 0.02, 0.0, 0, 0, 0, 1.0])
    rc_ref, x_ref, _, counts, _ = O.scan2scan(P, lc, lcr, ls, lsr, cs, cf, init)
    rc, x, st = eng.scan2scan(to_pcl(lc, lcr), to_pcl(ls, lsr), cs, cf, init)
    assert rc_ref == 1 and rc == 1
    assert np.array_equal(x, init) and np.array_equal(x_ref, init)
    assert st["status"] == 1 and st["n_edge"][0] == counts[0, 0] and st["n_plane"][0] == counts[0, 1]
    m = OdometryScanMatcher(eng)
    ok, pose = m.MatchScan2Scan(
        TimestampedPointCloud(cloud_corner_less_sharp=lc, ring_corner_less_sharp=lcr,
                              cloud_surf_less_flat=ls, ring_surf_less_flat=lsr),
        TimestampedPointCloud(cloud_corner_sharp=cs, cloud_surf_flat=cf), init)
    assert ok is False and np.array_equal(pose, init)


def _pair_batch():
    P, f0, f1, gt = _scan_pair()
    lc, lcr = f0["full"][f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]]
    ls, lsr = f0["full"][f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]]
    cs, cf = f1["full"][f1["idx_sharp"]], f1["full"][f1["idx_flat"]]
    lc1, lcr1 = f1["full"][f1["idx_less_sharp"]], f1["ring"][f1["idx_less_sharp"]]
    ls1, lsr1 = f1["full"][f1["idx_less_flat"]], f1["ring"][f1["idx_less_flat"]]
    cs0, cf0 = f0["full"][f0["idx_sharp"]], f0["full"][f0["idx_flat"]]
    empty = np.zeros((0, 4), np.float32)
    shifted = lc.copy()
    shifted[:, :3] += np.float32([40.0, -25.0, 3.0])  # another grid origin / extent for this pair
    sls = ls.copy()
    sls[:, :3] += np.float32([40.0, -25.0, 3.0])
    scs, scf = cs.copy(), cf.copy()
    scs[:, :3] += np.float32([40.0, -25.0, 3.0])
    scf[:, :3] += np.float32([40.0, -25.0, 3.0])
    pairs = [  # (last corner, rings, last surf, rings, sharp, flat, initial guess)
        (lc, lcr, ls, lsr, cs, cf, S.pose_identity()),
        (lc, lcr, ls, lsr, cs, cf, gt),
        (lc1, lcr1, ls1, lsr1, cs0, cf0, S.pose_inv(gt)),                            # the pair backwards
        (lc, lcr, ls, lsr, cs[:3], cf[:4], np.array([0.01, 0.02, 0.0, 0, 0, 0, 1.0])),  # too few
        (empty, np.zeros(0, np.uint16), ls, lsr, cs, cf, S.pose_identity()),            # nothing to search in
        (shifted, lcr, sls, lsr, scs, scf, S.pose_identity()),
        (lc, lcr, ls, lsr, empty, empty, gt),                                           # no queries
    ]
    return pairs


@pytest.mark.parametrize("budget", [None, 1])
def test_scan2scan_batch_equals_single_calls(budget, monkeypatch):
    """msfl_scan2scan_batch = B msfl_scan2scan calls (poses, statuses, correspondence counts, LM traces), with all
    pair grids indexed at once (budget None) and one pair per chunk (budget 1 cell -> every pair is its own chunk)."""
    if budget is not None:
        monkeypatch.setenv("MSFL_PAIR_CELL_BUDGET", str(budget))
    e = Engine()
    pairs = _pair_batch()
    single = [e.scan2scan(to_pcl(p[0], p[1]), to_pcl(p[2], p[3]), p[4], p[5], p[6]) for p in pairs]
    status, poses, st = e.scan2scan_batch([to_pcl(p[0], p[1]) for p in pairs], [to_pcl(p[2], p[3]) for p in pairs],
                                          [p[4] for p in pairs], [p[5] for p in pairs], np.stack([p[6] for p in pairs]),
                                          want_stats=True)
    assert list(status) == [s[0] for s in single] == [0, 0, 0, 1, 1, 0, 1]
    for b, (rc, x, s1) in enumerate(single):
        assert np.array_equal(poses[b], x), b
        assert st[b]["status"] == s1["status"] and st[b]["n_outer"] == s1["n_outer"], b
        assert st[b]["n_edge"] == s1["n_edge"] and st[b]["n_plane"] == s1["n_plane"], b
        for la, lb in zip(st[b]["lm"], s1["lm"]):
            assert la["n_attempts"] == lb["n_attempts"] and la["termination"] == lb["termination"]
    # and against the oracle for the first pair
    P = O.default_params()
    p = pairs[0]
    rc_ref, x_ref, _, _, _ = O.scan2scan(P, p[0], p[1], p[2], p[3], p[4], p[5], p[6])
    dt, dr = S.pose_error(poses[0], x_ref)
    assert rc_ref == 0 and dt < 1e-8 and dr < 1e-8


def test_scan2scan_rejects_unsorted_rings(eng):
    P, f0, f1, gt = _scan_pair()
    lc, lcr = f0["full"][f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]].copy()
    lcr[10], lcr[400] = lcr[400], lcr[10]
    ls, lsr = f0["full"][f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]]
    with pytest.raises(MsflError):
        eng.scan2scan(to_pcl(lc, lcr), to_pcl(ls, lsr), f1["full"][f1["idx_sharp"]], f1["full"][f1["idx_flat"]],
                      S.pose_identity())


def test_registration_to_mapping_pipeline(eng):
    """extract -> scan-to-scan through the reference-shaped host classes."""
    sc = S.make_scene()
    traj = S.trajectory(3)
    reg = ScanRegistration(eng)
    scans = [reg.extract(*S.raycast_scan(sc, "vlp16", traj[k], seed=60 + k)) for k in range(2)]
    ok, pose = OdometryScanMatcher(eng).MatchScan2Scan(scans[0], scans[1], S.pose_identity())
    gt = S.pose_mul(S.pose_inv(traj[0]), traj[1])
    dt, dr = S.pose_error(pose, gt)
    assert ok and dt < 0.05 and dr < 0.01


def test_batched_extraction_is_bitwise_the_single_scan_extraction(eng):
    """msfl_extract_features_batch: scans of different sensors and sizes in ONE launch sequence (grid.y = scan, one sort
    over (scan, ring)) give exactly what B single calls give -- and what the oracle gives."""
    scene40, scene80 = S.make_scene("room40"), S.make_scene("room80")
    traj = S.trajectory(4)
    raws = [S.raycast_scan(scene40, "vlp16", traj[0], seed=900), S.raycast_scan(scene80, "hdl64", traj[1], seed=901),
            S.raycast_scan(scene40, "vlp16", traj[2], seed=902), S.raycast_scan(scene80, "os1-128", traj[3], seed=903),
            S.raycast_scan(scene40, "vlp16", traj[1], seed=904)]
    T = np.array([0.1, -0.2, 0.3, 0.0, 0.0, np.sin(0.05), np.cos(0.05)])
    batch = eng.extract_features_batch([r[0] for r in raws], [r[1] for r in raws], T)
    P = O.default_params()
    for (xyzi, ring), fb in zip(raws, batch):
        fs = eng.extract_features(xyzi, ring, T)
        for k in ("full", "ring", "curvature", "label", "idx_sharp", "idx_less_sharp", "idx_flat", "idx_less_flat"):
            assert np.array_equal(fb[k], fs[k]), k
        _check_features(fb, O.extract_features(P, xyzi, ring, T))


def test_raw_to_pose_chain_equals_the_stage_by_stage_calls(eng, vlp16_case):
    """msfl_register_and_match_batch (registration + VoxelGrid x 2 + scan-to-map, everything batched on the GPU) against
    the same stages called one by one through the C ABI: bit-identical poses, same query counts; and against the oracle."""
    case = vlp16_case
    eng.set_submap(case["map_corner"], case["map_surf"])
    raws = [q["raw"] for q in case["queries"]] * 2   # 6 scans
    inits = [q["init"] for q in case["queries"]] * 2
    poses, counts = eng.register_and_match_batch(([r[0] for r in raws], [r[1] for r in raws]), inits, want_counts=True)
    corners, surfs = [], []
    for (xyzi, ring), c in zip(raws, counts):
        f = eng.extract_features(xyzi, ring, None)
        corners.append(eng.voxel_grid(f["full"][f["idx_less_sharp"]], 0.2))
        surfs.append(eng.voxel_grid(f["full"][f["idx_less_flat"]], 0.4))
        assert c["n_full"] == f["full"].shape[0] and c["n_less_sharp"] == len(f["idx_less_sharp"]) and c["n_less_flat"] == len(f["idx_less_flat"])
        assert c["n_corner_queries"] == corners[-1].shape[0] and c["n_surf_queries"] == surfs[-1].shape[0]
    rc, ref, _ = eng.scan2map_batch(corners, surfs, inits)
    assert rc == 0 and np.array_equal(poses, ref)
    P = O.default_params()
    for i, q in enumerate(case["queries"]):
        x_ref, _, _ = O.scan2map(P, case["map_corner"], case["map_surf"], q["corner"], q["surf"], q["init"])
        dt, dr = S.pose_error(poses[i], x_ref)
        # the oracle's queries carry the oracle's relative times in `intensity` (1-ulp atan2f difference); xyz are equal
        assert dt < 1e-8 and dr < 1e-8


def test_replay_batch_odometry_and_map_matching_from_one_registration_pass(eng, vlp16_case):
    """msfl_replay_batch on three consecutive raw scans: the odometry of the two pairs is bit-identical to
    msfl_scan2scan on the separately extracted features; with compose the map matcher starts from the dead-reckoned
    guesses pose[b-1] * odom[b] (laser_odometry.cc:79); without, it is msfl_register_and_match_batch."""
    case = vlp16_case
    eng.set_submap(case["map_corner"], case["map_surf"])
    qs = case["queries"]
    raws = ([q["raw"][0] for q in qs], [q["raw"][1] for q in qs])
    B = len(qs)
    ident = np.tile(S.pose_identity(), (B, 1))
    inits = np.stack([q["init"] for q in qs])
    od, status, poses = eng.replay_batch(raws, ident, inits, compose=True)
    feats = [eng.extract_features(x, r, None) for x, r in zip(*raws)]
    assert list(status) == [0] * B and np.array_equal(od[0], S.pose_identity())
    for b in range(1, B):
        f0, f1 = feats[b - 1], feats[b]
        rc, x, _ = eng.scan2scan(to_pcl(f0["full"][f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]]),
                                 to_pcl(f0["full"][f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]]),
                                 f1["full"][f1["idx_sharp"]], f1["full"][f1["idx_flat"]], S.pose_identity())
        assert rc == 0 and np.array_equal(od[b], x), b
        dt, dr = S.pose_error(od[b], S.pose_mul(S.pose_inv(qs[b - 1]["gt"]), qs[b]["gt"]))
        assert dt < 0.05 and dr < 0.01
    guesses = [inits[0]]
    for b in range(1, B):
        guesses.append(S.pose_mul(guesses[-1], od[b]))
    ref = eng.register_and_match_batch(raws, np.stack(guesses))
    for b in range(B):
        dt, dr = S.pose_error(poses[b], ref[b])
        assert dt < 1e-9 and dr < 1e-9, b
        dt, dr = S.pose_error(poses[b], qs[b]["gt"])
        assert dt < 0.05 and dr < 0.01, b
    od2, status2, poses2 = eng.replay_batch(raws, ident, inits, compose=False)
    assert np.array_equal(od2, od) and np.array_equal(poses2, eng.register_and_match_batch(raws, inits))
