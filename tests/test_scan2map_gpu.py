"""Parity of the CUDA scan-to-map path (through the C ABI) against the CPU oracle.

Tolerances: kNN indices bit-exact (integer/fp32 compare-select work); line/plane fit constants
1e-9 absolute (fp64, different FMA contraction); pose <= 1e-4 m / 1e-4 rad (north_star) -- in
practice ~1e-10; LM trace (attempts, accept/reject pattern) identical, costs 1e-9 relative.
"""
import numpy as np
import pytest

import oracle as O
from msf_loam_b200 import Engine, MappingScanMatcher, TimestampedPointCloud, default_params, to_pcl
from msf_loam_b200 import synth as S

pytestmark = pytest.mark.gpu

POSE_TOL_M = 1e-4
POSE_TOL_RAD = 1e-4


@pytest.fixture(scope="module")
def eng(vlp16_case):
    e = Engine()
    e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
    yield e
    e.close()


def _oracle_corr_full(P, case, q, pose):
    """oracle association expanded to one row per query (zeros where no factor)."""
    corr, ne, npl, kidx = O.associate_map(P, case["map_corner"], case["map_surf"], q["corner"], q["surf"], pose)
    return corr, ne, npl, kidx


def test_knn_indices_bit_exact(eng, vlp16_case):
    P = O.default_params()
    for q in vlp16_case["queries"]:
        knn, _ = eng.associate_map(q["corner"], q["surf"], q["init"])
        _, _, _, kidx = _oracle_corr_full(P, vlp16_case, q, q["init"])
        assert knn.shape == kidx.shape
        assert np.array_equal(knn, kidx)
        assert (knn[:, 0] >= 0).sum() > 1000  # the case is not degenerate


def test_fit_constants_match(eng, vlp16_case):
    P = O.default_params()
    q = vlp16_case["queries"][0]
    _, corr_gpu = eng.associate_map(q["corner"], q["surf"], q["init"])
    corr, ne, npl, _ = _oracle_corr_full(P, vlp16_case, q, q["init"])
    has = np.any(corr_gpu[:, 3:] != 0, axis=1)
    nc = q["corner"].shape[0]
    assert has[:nc].sum() == ne and has[nc:].sum() == npl
    got = corr_gpu[has]
    assert np.abs(got[:, :3] - corr[:, 4:7]).max() < 1e-9
    dn = np.minimum(np.abs(got[:, 3:] - corr[:, 7:10]).max(axis=1), np.abs(got[:, 3:] + corr[:, 7:10]).max(axis=1))
    assert dn.max() < 1e-9


def test_accumulate_matches_oracle(eng, vlp16_case):
    P = O.default_params()
    q = vlp16_case["queries"][1]
    corr, ne, npl, _ = _oracle_corr_full(P, vlp16_case, q, q["init"])
    cost, H, g = O.accumulate(P, corr, q["init"])
    cost_g, H_g, g_g = eng.accumulate(corr[:, 1:4], corr[:, 4:10], ne, npl, q["init"])
    assert abs(cost_g - cost) <= 1e-12 * abs(cost)
    assert np.abs(H_g - H).max() <= 1e-11 * np.abs(H).max()
    assert np.abs(g_g - g).max() <= 1e-11 * np.abs(g).max()


@pytest.mark.parametrize("schedule", ["reference", "fixed10"])
def test_pose_parity(vlp16_case, schedule):
    over = {} if schedule == "reference" else {"early_exit": 0, "max_num_iterations": 5}
    P = O.default_params(**over)
    e = Engine(default_params(**over))
    e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
    for q in vlp16_case["queries"]:
        x_ref, logs, counts = O.scan2map(P, vlp16_case["map_corner"], vlp16_case["map_surf"], q["corner"], q["surf"], q["init"])
        rc, x, st = e.scan2map(q["corner"], q["surf"], q["init"])
        assert rc == 0
        dt, dr = S.pose_error(x, x_ref)
        assert dt <= POSE_TOL_M and dr <= POSE_TOL_RAD
        assert dt < 1e-8 and dr < 1e-8  # what the fp64 path actually achieves
        assert st["n_edge"] == list(counts[:, 0]) and st["n_plane"] == list(counts[:, 1])
        for lg, lr in zip(st["lm"], logs):
            assert lg["n_attempts"] == lr["n_attempts"] and lg["termination"] == lr["termination"]
            assert [i["accepted"] for i in lg["iters"]] == [i["accepted"] for i in lr["iters"]]
            assert abs(lg["final_cost"] - lr["final_cost"]) <= 1e-9 * lr["final_cost"]
        # and the estimate is a sane localisation (synthetic noise 1 cm)
        dt_gt, dr_gt = S.pose_error(x, q["gt"])
        assert dt_gt < 0.03 and dr_gt < 0.005
    e.close()


def test_batch_equals_singles_bitwise(eng, vlp16_case):
    qs = vlp16_case["queries"]
    singles = [eng.scan2map(q["corner"], q["surf"], q["init"])[1] for q in qs]
    # ragged batch incl. an empty scan and a repeated one
    empty = np.zeros((0, 4), np.float32)
    corners = [qs[0]["corner"], empty, qs[1]["corner"], qs[2]["corner"], qs[0]["corner"]]
    surfs = [qs[0]["surf"], empty, qs[1]["surf"], qs[2]["surf"], qs[0]["surf"]]
    inits = [qs[0]["init"], qs[1]["init"], qs[1]["init"], qs[2]["init"], qs[0]["init"]]
    rc, xs, st = eng.scan2map_batch(corners, surfs, inits, want_stats=True)
    assert rc == 0
    assert np.array_equal(xs[0], singles[0]) and np.array_equal(xs[2], singles[1])
    assert np.array_equal(xs[3], singles[2]) and np.array_equal(xs[4], singles[0])
    assert np.array_equal(xs[1], np.asarray(inits[1]))  # empty scan: pose untouched
    assert st[1]["n_edge"] == [0, 0] and st[1]["n_plane"] == [0, 0]


def test_sorted_association_path_is_bitwise_identical(vlp16_case):
    """assoc_sorted = 2 (queries ordered by submap cell, used for large batches) vs 1 (flat order)."""
    P = O.default_params()
    qs = vlp16_case["queries"]
    res = {}
    for mode in (1, 2, 3):  # 3: cell order + TMA-staged shared-memory tiles (k_knn5_tiled)
        e = Engine(default_params(assoc_sorted=mode))
        e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
        knn, corr = e.associate_map(qs[0]["corner"], qs[0]["surf"], qs[0]["init"])
        far = np.array([500.0, 500.0, 50.0, 0, 0, 0, 1.0])
        rc, xs, st = e.scan2map_batch([q["corner"] for q in qs] + [qs[0]["corner"]],
                                      [q["surf"] for q in qs] + [qs[0]["surf"]],
                                      [q["init"] for q in qs] + [far], want_stats=True)
        res[mode] = (knn, corr, xs, [s["n_edge"] + s["n_plane"] for s in st])
        e.close()
    for mode in (2, 3):
        assert np.array_equal(res[1][0], res[mode][0]) and np.array_equal(res[1][1], res[mode][1])
        assert np.array_equal(res[1][2], res[mode][2]) and res[1][3] == res[mode][3]
    _, _, _, kidx = _oracle_corr_full(P, vlp16_case, qs[0], qs[0]["init"])
    assert np.array_equal(res[2][0], kidx)


def test_large_host_batch_pipelined_path_bitwise(eng, vlp16_case):
    """B = 512 host batch (> 2M queries) takes the chunked H2D/compute pipeline; every replica must
    equal the single-scan result bit for bit, for packed-contiguous and for PCL-layout inputs."""
    qs = vlp16_case["queries"]
    singles = [eng.scan2map(q["corner"], q["surf"], q["init"])[1] for q in qs]
    B = 512
    inits = np.stack([qs[i % 3]["init"] for i in range(B)])
    # (a) one contiguous pinned-style buffer per class
    cat_c = np.concatenate([qs[i % 3]["corner"] for i in range(B)])
    cat_s = np.concatenate([qs[i % 3]["surf"] for i in range(B)])
    co = np.concatenate([[0], np.cumsum([qs[i % 3]["corner"].shape[0] for i in range(B)])])
    so = np.concatenate([[0], np.cumsum([qs[i % 3]["surf"].shape[0] for i in range(B)])])
    rc, xs, _ = eng.scan2map_batch([cat_c[co[i]:co[i + 1]] for i in range(B)], [cat_s[so[i]:so[i + 1]] for i in range(B)], inits)
    assert rc == 0
    for i in range(B):
        assert np.array_equal(xs[i], singles[i % 3])
    # (b) separate 32-byte PCL-layout clouds (host repack per chunk)
    pc = [to_pcl(q["corner"]) for q in qs]
    ps = [to_pcl(q["surf"]) for q in qs]
    rc, xs2, st = eng.scan2map_batch([pc[i % 3] for i in range(B)], [ps[i % 3] for i in range(B)], inits, want_stats=True)
    assert rc == 0 and np.array_equal(xs2, xs)
    assert st[B - 1]["n_plane"] == st[(B - 1) % 3]["n_plane"] and st[B - 1]["lm"][1]["n_attempts"] > 0


@pytest.mark.parametrize("G", [2, 4, 8, 16])
def test_cluster_lm_matches_oracle(vlp16_case, G):
    """lm_cluster = G: one thread-block cluster of G CTAs per scan, partial sums combined over DSMEM.  A single scan
    (e.scan2map) takes the fused one-launch kernel (scan2map_fused.cu: association into shared memory + solve), a batch
    the clustered LM kernel.  The summation grouping differs from G = 1, so parity is to rounding (1e-8), not bitwise."""
    P = O.default_params()
    e = Engine(default_params(lm_cluster=G))
    e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
    qs = vlp16_case["queries"]
    for q in qs:
        x_ref, logs, counts = O.scan2map(P, vlp16_case["map_corner"], vlp16_case["map_surf"], q["corner"], q["surf"], q["init"])
        rc, x, st = e.scan2map(q["corner"], q["surf"], q["init"])
        dt, dr = S.pose_error(x, x_ref)
        assert rc == 0 and dt < 1e-8 and dr < 1e-8
        assert st["n_edge"] == list(counts[:, 0]) and st["n_plane"] == list(counts[:, 1])
        assert [l["n_attempts"] for l in st["lm"]] == [l["n_attempts"] for l in logs]
    # ragged batch with an empty scan and a no-correspondence scan through the clustered kernel
    empty = np.zeros((0, 4), np.float32)
    far = np.array([500.0, 500.0, 50.0, 0, 0, 0, 1.0])
    rc, xs, st = e.scan2map_batch([qs[0]["corner"], empty, qs[1]["corner"]], [qs[0]["surf"], empty, qs[1]["surf"]],
                                  [qs[0]["init"], qs[1]["init"], far], want_stats=True)
    assert rc == 0 and np.array_equal(xs[1], qs[1]["init"]) and np.array_equal(xs[2], far)
    assert S.pose_error(xs[0], qs[0]["gt"])[0] < 0.03
    # fused kernel: no correspondence at all -> pose untouched, zero counts; and both LM schedules
    rc, x, st = e.scan2map(qs[0]["corner"], qs[0]["surf"], far)
    assert rc == 0 and np.array_equal(x, far) and st["n_edge"] == [0, 0] and st["n_plane"] == [0, 0]
    e.close()
    e = Engine(default_params(lm_cluster=G, early_exit=0, max_num_iterations=5))
    e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
    Pf = O.default_params(early_exit=0, max_num_iterations=5)
    x_ref, logs, _ = O.scan2map(Pf, vlp16_case["map_corner"], vlp16_case["map_surf"], qs[2]["corner"], qs[2]["surf"], qs[2]["init"])
    rc, x, st = e.scan2map(qs[2]["corner"], qs[2]["surf"], qs[2]["init"])
    dt, dr = S.pose_error(x, x_ref)
    assert rc == 0 and dt < 1e-8 and dr < 1e-8
    for l, lr in zip(st["lm"], logs):
        assert [it["accepted"] for it in l["iters"]] == [it["accepted"] for it in lr["iters"]]
        assert abs(l["final_cost"] - lr["final_cost"]) <= 1e-9 * lr["final_cost"]
    e.close()


def test_no_correspondence_leaves_pose_untouched(eng, vlp16_case):
    q = vlp16_case["queries"][0]
    far = np.array([500.0, 500.0, 50.0, 0, 0, 0, 1.0])
    rc, x, st = eng.scan2map(q["corner"], q["surf"], far)
    assert rc == 0 and np.array_equal(x, far)
    assert st["n_edge"] == [0, 0] and st["n_plane"] == [0, 0]
    x_ref, _, _ = O.scan2map(O.default_params(), vlp16_case["map_corner"], vlp16_case["map_surf"], q["corner"], q["surf"], far)
    assert np.array_equal(x_ref, far)


def test_pcl_point_layout_and_matcher_interface(vlp16_case):
    """Same call shape as the reference: MatchScan2Map(cloud_map, scan_curr, false, ..., &pose) with
    32-byte pcl::PointXYZI clouds (stride/offset marshalling)."""
    q = vlp16_case["queries"][0]
    m = MappingScanMatcher()
    cloud_map = TimestampedPointCloud(cloud_corner_less_sharp=to_pcl(vlp16_case["map_corner"]),
                                      cloud_surf_less_flat=to_pcl(vlp16_case["map_surf"]))
    scan = TimestampedPointCloud(cloud_corner_less_sharp=to_pcl(q["corner"]), cloud_surf_less_flat=to_pcl(q["surf"]))
    ok, pose = m.MatchScan2Map(cloud_map, scan, False, q["init"])
    assert ok is True
    x_ref, _, _ = O.scan2map(O.default_params(), vlp16_case["map_corner"], vlp16_case["map_surf"], q["corner"], q["surf"], q["init"])
    dt, dr = S.pose_error(pose, x_ref)
    assert dt < 1e-8 and dr < 1e-8
    m.engine.close()


def test_requires_submap():
    from msf_loam_b200 import MsflError
    e = Engine()
    with pytest.raises(MsflError):
        e.scan2map(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32), S.pose_identity())
    e.close()


def test_radix_sort_fallback_of_the_cell_order_is_bitwise_identical(vlp16_case, monkeypatch):
    """Large batches order their queries by submap cell: counting sort while the bin table is small, cub radix sort
    for far-spread submaps (forced here through MSFL_COUNT_SORT_MAX_BINS).  The order is only a locality hint."""
    qs = vlp16_case["queries"]
    B = 40
    res = []
    for max_bins in ("0", None):
        if max_bins is None:
            monkeypatch.delenv("MSFL_COUNT_SORT_MAX_BINS", raising=False)
        else:
            monkeypatch.setenv("MSFL_COUNT_SORT_MAX_BINS", max_bins)
        e = Engine(default_params(assoc_sorted=2))
        e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
        rc, xs, st = e.scan2map_batch([qs[i % 3]["corner"] for i in range(B)], [qs[i % 3]["surf"] for i in range(B)],
                                      [qs[i % 3]["init"] for i in range(B)], want_stats=True)
        res.append((xs, [s["n_edge"] + s["n_plane"] for s in st]))
        e.close()
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
    assert np.array_equal(res[0][0][0], res[0][0][3])  # replicas of one scan agree


def test_tma_staged_tile_search_matches_direct_search_on_a_dense_batch(vlp16_case):
    """assoc_sorted = 3 (k_knn5_tiled: the CTA's 3x3x3 neighbourhood staged in shared memory by cp.async.bulk) on a
    batch dense enough that most CTAs sit inside one cell, incl. perturbed replicas that straddle cell borders."""
    qs = vlp16_case["queries"]
    rng = np.random.default_rng(5)
    B = 96
    inits = np.stack([S.perturb_pose(qs[i % 3]["gt"], rng) for i in range(B)])
    out = {}
    for mode in (2, 3):
        e = Engine(default_params(assoc_sorted=mode))
        e.set_submap(vlp16_case["map_corner"], vlp16_case["map_surf"])
        rc, xs, st = e.scan2map_batch([qs[i % 3]["corner"] for i in range(B)], [qs[i % 3]["surf"] for i in range(B)], inits,
                                      want_stats=True)
        out[mode] = (xs, [s["n_edge"] + s["n_plane"] for s in st])
        e.close()
    assert np.array_equal(out[2][0], out[3][0]) and out[2][1] == out[3][1]
