"""Edge cases of the CUDA path against the oracle: non-finite queries, exact distance ties (duplicated map
points), a map far from the origin / at negative coordinates, tiny maps, degenerate neighbourhoods."""
import numpy as np
import pytest

import oracle as O
from msf_loam_b200 import Engine, MsflError, default_params
from msf_loam_b200 import synth as S

pytestmark = pytest.mark.gpu


def _shifted_case(case, shift):
    sh = np.array(list(shift) + [0], np.float32)
    T = np.array([shift[0], shift[1], shift[2], 0, 0, 0, 1.0])
    qs = []
    for q in case["queries"]:
        qs.append({"corner": q["corner"], "surf": q["surf"], "init": S.pose_mul(T, q["init"]), "gt": S.pose_mul(T, q["gt"])})
    return {"map_corner": case["map_corner"] + sh, "map_surf": case["map_surf"] + sh, "queries": qs}


@pytest.mark.parametrize("shift", [(-3000.25, 4100.5, -120.75), (0.5, 0.5, 0.5)])
def test_map_far_from_origin_and_negative_coordinates(vlp16_case, shift):
    c = _shifted_case(vlp16_case, shift)
    P = O.default_params()
    for mode in (1, 2):
        e = Engine(default_params(assoc_sorted=mode))
        e.set_submap(c["map_corner"], c["map_surf"])
        q = c["queries"][0]
        knn, _ = e.associate_map(q["corner"], q["surf"], q["init"])
        _, _, _, kidx = O.associate_map(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
        assert np.array_equal(knn, kidx) and (knn[:, 0] >= 0).sum() > 1000
        x_ref, _, _ = O.scan2map(P, c["map_corner"], c["map_surf"], q["corner"], q["surf"], q["init"])
        rc, x, _ = e.scan2map(q["corner"], q["surf"], q["init"])
        dt, dr = S.pose_error(x, x_ref)
        assert rc == 0 and dt < 1e-7 and dr < 1e-9
        e.close()


def test_duplicate_map_points_tie_break_on_index(vlp16_case):
    """every surf map point stored three times: 5-NN sets are full of exact distance ties."""
    c = vlp16_case
    mc = c["map_corner"]
    ms = np.concatenate([c["map_surf"], c["map_surf"][::-1], c["map_surf"]])
    e = Engine()
    e.set_submap(mc, ms)
    q = c["queries"][1]
    knn, _ = e.associate_map(q["corner"], q["surf"], q["init"])
    _, _, _, kidx = O.associate_map(O.default_params(), mc, ms, q["corner"], q["surf"], q["init"])
    assert np.array_equal(knn, kidx)
    e.close()


def test_non_finite_and_far_queries_create_no_factor(vlp16_case):
    c = vlp16_case
    q = c["queries"][0]
    surf = q["surf"].copy()
    surf[5, 0] = np.nan
    surf[6, 1] = np.inf
    surf[7, :3] = 1e30
    surf[8, :3] = -1e9
    e = Engine()
    e.set_submap(c["map_corner"], c["map_surf"])
    knn, corr = e.associate_map(q["corner"], surf, q["init"])
    nc = q["corner"].shape[0]
    assert np.all(knn[nc + 5:nc + 9] == -1) and np.all(corr[nc + 5:nc + 9] == 0)
    rc, x, st = e.scan2map(q["corner"], surf, q["init"])
    assert rc == 0 and np.all(np.isfinite(x)) and S.pose_error(x, q["gt"])[0] < 0.03
    e.close()


def test_tiny_and_degenerate_maps(vlp16_case):
    c = vlp16_case
    q = c["queries"][0]
    P = O.default_params()
    # fewer than 5 points per class: no neighbourhood can pass the gate
    e = Engine()
    e.set_submap(c["map_corner"][:3], c["map_surf"][:4])
    rc, x, st = e.scan2map(q["corner"], q["surf"], q["init"])
    assert rc == 0 and np.array_equal(x, q["init"]) and st["n_plane"] == [0, 0]
    # all map points on one line (rank-deficient plane fits, zero-variance directions)
    line = np.zeros((200, 4), np.float32)
    line[:, 0] = np.linspace(-5, 5, 200)
    e.set_submap(line, line)
    pts = np.zeros((50, 4), np.float32)
    pts[:, 0] = np.linspace(-4, 4, 50); pts[:, 1] = 0.05
    ident = S.pose_identity()
    knn, corr = e.associate_map(pts, pts, ident)
    c_ref, ne, npl, kidx = O.associate_map(P, line, line, pts, pts, ident)
    assert np.array_equal(knn, kidx)
    has = np.any(corr[:, 3:] != 0, axis=1)
    assert has[:50].sum() == ne and np.all(np.isfinite(corr))
    with pytest.raises(MsflError):
        e.set_submap(np.zeros((0, 4), np.float32), line)
    bad = line.copy(); bad[3, 2] = np.nan
    with pytest.raises(MsflError):
        e.set_submap(line, bad)
    e.close()


def test_far_spread_submap_large_batch_uses_reduced_subcell_keys(vlp16_case):
    """A submap whose two classes span > 2^26 cells in total (two sites 3.2 km apart): the cell-ordered association used
    to refuse it (cell id + 6 sub-cell bits > 32 bits) once a batch reached 65536 queries, although single scans worked.
    It now drops to 2x2x2 sub-cells (radix-sorted keys): same neighbours as the flat path, and a >= 65536-query batch runs."""
    far = np.array([2500.0, 1990.0, 0.0, 0.0], np.float32)
    mc = np.concatenate([vlp16_case["map_corner"], vlp16_case["map_corner"] + far])
    ms = np.concatenate([vlp16_case["map_surf"], vlp16_case["map_surf"] + far])
    qs = vlp16_case["queries"]
    P = O.default_params()
    e = Engine(default_params())  # auto mode
    e.set_submap(mc, ms)
    q = qs[0]
    flat = Engine(default_params(assoc_sorted=1))
    flat.set_submap(mc, ms)
    knn_flat, corr_flat = flat.associate_map(q["corner"], q["surf"], q["init"])
    srt = Engine(default_params(assoc_sorted=2))
    srt.set_submap(mc, ms)
    knn_sorted, corr_sorted = srt.associate_map(q["corner"], q["surf"], q["init"])
    assert np.array_equal(knn_flat, knn_sorted) and np.array_equal(corr_flat, corr_sorted)
    _, _, _, kidx = O.associate_map(P, mc, ms, q["corner"], q["surf"], q["init"])
    assert np.array_equal(knn_sorted, kidx)
    B = 16  # 16 x ~4.7 k queries >= 65536
    n_q = sum(qs[i % 3]["corner"].shape[0] + qs[i % 3]["surf"].shape[0] for i in range(B))
    assert n_q >= 65536
    rc, xs, _ = e.scan2map_batch([qs[i % 3]["corner"] for i in range(B)], [qs[i % 3]["surf"] for i in range(B)],
                                 [qs[i % 3]["init"] for i in range(B)])
    assert rc == 0
    for i in range(3):
        x_ref, _, _ = O.scan2map(P, mc, ms, qs[i]["corner"], qs[i]["surf"], qs[i]["init"])
        dt, dr = S.pose_error(xs[i], x_ref)
        assert dt < 1e-7 and dr < 1e-9
        assert np.array_equal(xs[i], xs[i + 3])
    for x in (e, flat, srt):
        x.close()
