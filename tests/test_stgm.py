"""SURVEY.md 8f row 1 -- GPU-resident STGM submap producer (HybridGrid) against the CPU oracle, and both against the
reference's own slam/map/hybrid_grid.cc compiled unmodified (oracle/_ref, oracle/ref_map_shim.cc).  The reference
concatenates the cells of a surround query in hash order of their shared pointers (heap addresses), so its clouds are
compared as multisets of points."""
import numpy as np
import pytest

import oracle as O
from msf_loam_b200 import synth as S


def _scans(n=4, seed0=500):
    P = O.default_params()
    scene, traj = S.make_scene(), S.trajectory(n)
    out = []
    for k in range(n):
        f = O.extract_features(P, *S.raycast_scan(scene, "vlp16", traj[k], seed=seed0 + k), None)
        out.append((f["full"][f["idx_less_sharp"]], f["full"][f["idx_less_flat"]], traj[k]))
    return out


def test_oracle_stgm_properties():
    scans = _scans(3)
    m = O.Stgm(3.0, 0.4)
    sizes = []
    for corner, surf, pose in scans:
        m.insert(S.transform_cloud(pose, surf))
        sizes.append(m.size())
    assert sizes[0][0] < sizes[1][0] < sizes[2][0] and sizes[2][1] >= sizes[0][1]
    allpts = m.dump()
    # every cell was voxel-filtered: inside a 3 m cell no two points share a 0.4 m voxel (a voxel that
    # straddles a cell border legitimately yields one centroid per cell)
    vox = np.floor(allpts[:, :3] * (np.float32(1.0) / np.float32(0.4))).astype(np.int64)
    cellid = np.round(allpts[:, :3] / np.float32(3.0)).astype(np.int64)
    assert len(np.unique(np.concatenate([cellid, vox], axis=1), axis=0)) >= 0.995 * len(allpts)
    # a single insert of a cloud == VoxelGrid per 3 m cell
    one = O.Stgm(3.0, 0.4)
    w = S.transform_cloud(scans[0][2], scans[0][1])
    one.insert(w)
    cell = np.round(w[:, :3].astype(np.float32) / np.float32(3.0)).astype(np.int64)
    expect = sum(len(O.voxel_grid(w[np.all(cell == c, axis=1)], 0.4)) for c in np.unique(cell, axis=0))
    assert one.size()[0] == expect
    # surround from inside the mapped area returns (almost) the whole map; from far away, nothing
    sur = m.surround(scans[2][1], scans[2][2])
    assert 0.9 * len(allpts) <= len(sur) <= len(allpts)
    far = np.array([500.0, 0, 0, 0, 0, 0, 1.0])
    assert len(m.surround(scans[2][1], far)) == 0


def _rows_sorted(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a[np.lexsort(a.T[::-1])]


def _ref_available():
    from oracle import ref as R
    return R.available()


@pytest.mark.skipif(not _ref_available(), reason="no reference checkout and no prebuilt oracle/_ref library")
def test_oracle_stgm_equals_the_reference_hybrid_grid():
    """Insert four scans (two feature classes, their leaf sizes), query after every insert: the oracle's surround cloud
    is the reference's, point for point (as a multiset); a far-away query returns nothing; points beyond 60 m are
    ignored by the query (hybrid_grid.cc:474)."""
    from oracle import ref as R
    scans = _scans(4)
    for cls, leaf in ((0, 0.2), (1, 0.4)):
        r, o = R.Map(3.0, leaf), O.Stgm(3.0, leaf)
        for k, sc in enumerate(scans):
            w = S.transform_cloud(sc[2], sc[cls])
            r.insert(w)
            o.insert(w)
            nxt = scans[(k + 1) % len(scans)]
            guess = S.perturb_pose(nxt[2], np.random.default_rng(k))
            a, b = r.surround(nxt[cls], guess), o.surround(nxt[cls], guess)
            assert a.shape == b.shape and a.shape[0] > 100
            assert np.array_equal(_rows_sorted(a), _rows_sorted(b))
        far = np.array([500.0, 0, 0, 0, 0, 0, 1.0])
        assert len(r.surround(scans[0][cls], far)) == 0
        beyond = scans[0][cls].copy()
        beyond[:, 0] += 100.0  # every point farther than kDist = 60 m from the sensor
        assert len(r.surround(beyond, scans[0][2])) == 0 and len(o.surround(beyond, scans[0][2])) == 0


@pytest.mark.gpu
@pytest.mark.skipif(not _ref_available(), reason="no reference checkout and no prebuilt oracle/_ref library")
def test_cuda_stgm_equals_the_reference_hybrid_grid():
    from msf_loam_b200 import Engine, HybridGrid
    from oracle import ref as R
    scans = _scans(4)
    e = Engine()
    try:
        for cls, leaf in ((0, 0.2), (1, 0.4)):
            r, g = R.Map(3.0, leaf), HybridGrid(e, 3.0, leaf)
            for k, sc in enumerate(scans):
                r.insert(S.transform_cloud(sc[2], sc[cls]))
                g.InsertScan(sc[cls], sc[2])
                nxt = scans[(k + 1) % len(scans)]
                guess = S.perturb_pose(nxt[2], np.random.default_rng(k))
                assert np.array_equal(_rows_sorted(r.surround(nxt[cls], guess)), _rows_sorted(g.GetSurroundedCloud(nxt[cls], guess)))
            g.close()
    finally:
        e.close()


@pytest.mark.gpu
def test_cuda_stgm_matches_oracle_bitwise_and_feeds_scan2map():
    from msf_loam_b200 import Engine, HybridGrid, set_submap_from_maps
    scans = _scans(4)
    e = Engine()
    gc, gs = HybridGrid(e, 3.0, 0.2), HybridGrid(e, 3.0, 0.4)
    oc, os_ = O.Stgm(3.0, 0.2), O.Stgm(3.0, 0.4)
    for k, (corner, surf, pose) in enumerate(scans[:3]):
        if k == 1:   # world-frame input path
            gc.InsertScan(S.transform_cloud(pose, corner)); gs.InsertScan(S.transform_cloud(pose, surf))
        else:        # sensor-frame scan + pose (TransformPointCloud on the device)
            gc.InsertScan(corner, pose); gs.InsertScan(surf, pose)
        oc.insert(S.transform_cloud(pose, corner)); os_.insert(S.transform_cloud(pose, surf))
        assert gc.size() == oc.size() and gs.size() == os_.size()
        assert np.array_equal(gc.dump(), oc.dump()) and np.array_equal(gs.dump(), os_.dump())
    corner, surf, pose = scans[3]
    guess = S.perturb_pose(pose, np.random.default_rng(1))
    sc, ss = gc.GetSurroundedCloud(corner, guess), gs.GetSurroundedCloud(surf, guess)
    assert np.array_equal(sc, oc.surround(corner, guess)) and np.array_equal(ss, os_.surround(surf, guess))
    assert len(sc) > 10 and len(ss) > 50  # the reference's gate (laser_mapping.cc:284-285)
    # the device-resident surround clouds become the submap without touching the host
    q_corner, q_surf = e.voxel_grid(corner, 0.2), e.voxel_grid(surf, 0.4)
    set_submap_from_maps(e, gc, gs)
    rc, x_dev, st = e.scan2map(q_corner, q_surf, guess)
    e.set_submap(sc, ss)
    rc2, x_host, st2 = e.scan2map(q_corner, q_surf, guess)
    assert rc == rc2 == 0 and np.array_equal(x_dev, x_host)
    x_ref, _, _ = O.scan2map(O.default_params(), sc, ss, q_corner, q_surf, guess)
    dt, dr = S.pose_error(x_dev, x_ref)
    assert dt < 1e-8 and dr < 1e-8
    dt, dr = S.pose_error(x_dev, pose)
    assert dt < 0.05 and dr < 0.01
    # empty map / empty scan edge cases
    empty = HybridGrid(e, 3.0, 0.4)
    assert len(empty.GetSurroundedCloud(surf, guess)) == 0
    empty.InsertScan(np.zeros((0, 4), np.float32))
    assert empty.size() == (0, 0)
    for g in (gc, gs, empty):
        g.close()
    e.close()


@pytest.mark.gpu
def test_mapping_frame_equals_the_separate_calls_and_the_reference_loop():
    """msfl_mapping_frame = one frame of LaserMapping (laser_mapping.cc:258-340).  Six frames along a trajectory, each
    starting from a perturbed pose: the frame call against (a) the same steps issued as separate C-ABI calls on a second
    pair of maps -- bitwise: poses, gate decisions, both maps -- and (b) the reference's own loop body assembled from its
    compiled HybridGrid and MatchScan2Map (oracle/_ref) with the oracle's VoxelGrid."""
    import time
    from msf_loam_b200 import Engine, HybridGrid, mapping_frame, set_submap_from_maps
    from oracle import ref as R
    scans = _scans(6)
    e = Engine()
    try:
        fc, fs = HybridGrid(e, 3.0, 0.2), HybridGrid(e, 3.0, 0.4)   # driven by the frame call
        sc, ss = HybridGrid(e, 3.0, 0.2), HybridGrid(e, 3.0, 0.4)   # driven by separate calls
        have_ref = R.available()
        rc_, rs_ = (R.Map(3.0, 0.2), R.Map(3.0, 0.4)) if have_ref else (None, None)
        t_frame = t_sep = 0.0
        for k, (corner, surf, gt) in enumerate(scans):
            guess = gt if k == 0 else S.perturb_pose(gt, np.random.default_rng(40 + k))
            t0 = time.perf_counter()
            matched, pose, st = mapping_frame(e, fc, fs, corner, surf, guess)
            t_frame += time.perf_counter() - t0
            # (a) the same frame as separate calls
            t0 = time.perf_counter()
            n_c, n_s = sc.GetSurroundedCloud(corner, guess, download=False), ss.GetSurroundedCloud(surf, guess, download=False)
            pose_sep, gate = guess.copy(), n_c > 10 and n_s > 50
            if gate:
                set_submap_from_maps(e, sc, ss)
                _, pose_sep, st_sep = e.scan2map(e.voxel_grid(corner, 0.2), e.voxel_grid(surf, 0.4), guess)
            sc.InsertScan(corner, pose_sep)
            ss.InsertScan(surf, pose_sep)
            t_sep += time.perf_counter() - t0
            assert matched == gate == (k > 0)  # the first frame only fills the empty maps
            assert np.array_equal(pose, pose_sep)
            if gate:
                assert st["n_edge"] == st_sep["n_edge"] and st["n_plane"] == st_sep["n_plane"] and st["n_plane"][0] > 1000
                dt, dr = S.pose_error(pose, gt)
                assert dt < 0.05 and dr < 0.01
            assert fc.size() == sc.size() and fs.size() == ss.size()
            assert np.array_equal(fc.dump(), sc.dump()) and np.array_equal(fs.dump(), ss.dump())
            # (b) the reference's loop body
            if have_ref:
                m_c, m_s = rc_.surround(corner, guess), rs_.surround(surf, guess)
                pose_ref = guess.copy()
                if len(m_c) > 10 and len(m_s) > 50:
                    ok, pose_ref = R.scan2map(m_c, m_s, O.voxel_grid(corner, 0.2), O.voxel_grid(surf, 0.4), guess)
                assert (len(m_c) > 10 and len(m_s) > 50) == matched
                dt, dr = S.pose_error(pose, pose_ref)
                assert dt < 1e-8 and dr < 1e-8, (k, dt, dr)
                # keep the reference maps on the same trajectory as ours (poses agree to ~1e-12, the maps to float rounding)
                rc_.insert(S.transform_cloud(pose, corner))
                rs_.insert(S.transform_cloud(pose, surf))
        print(f"mapping frame: {t_frame / len(scans) * 1e3:.2f} ms per frame in one call, {t_sep / len(scans) * 1e3:.2f} ms as separate calls")
        for g in (fc, fs, sc, ss):
            g.close()
    finally:
        e.close()
