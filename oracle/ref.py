"""ctypes bindings over oracle/_ref/libmsfl_ref.so -- the REFERENCE's own hot-path sources (scan registration, both
scan matchers, factors, parameterisation, GetDeltaQP, HybridGrid) compiled unmodified (oracle/ref_shim.cc,
ref_extract_shim.cc and ref_map_shim.cc say what is the reference's and what is stood in).  TEST INFRASTRUCTURE ONLY, like the rest of
oracle/: only tests/ may import it.  `available()` is False when there is neither a reference checkout to compile nor a
prebuilt library (the GPU box gets the prebuilt file with the repo snapshot)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import LmLog, _f32, _pose, _ptr, ref_lib

_D = C.POINTER(C.c_double)


def available() -> bool:
    return ref_lib() is not None


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def reset_logs(record_knn=False):
    ref_lib().msflref_log_reset(C.c_int(1 if record_knn else 0))


def solves():
    """One dict per ceres::Solve call since the last reset: supported, n_edge, n_plane, lm (the iteration log)."""
    L = ref_lib()
    out = []
    for i in range(L.msflref_n_solves()):
        sup, ne, npl, log = C.c_int(), C.c_int(), C.c_int(), LmLog()
        L.msflref_solve_info(C.c_int(i), C.byref(sup), C.byref(ne), C.byref(npl), C.byref(log))
        out.append({"supported": bool(sup.value), "n_edge": ne.value, "n_plane": npl.value, "lm": log.as_dict()})
    return out


def knn_log():
    """Every nearestKSearch the reference issued since the last reset(record_knn=True): (k_of[n], idx[list of arrays], d2)."""
    L = ref_lib()
    n, m = L.msflref_knn_log_searches(), L.msflref_knn_log_entries()
    k_of, idx, d2 = np.zeros(n, np.int32), np.zeros(m, np.int32), np.zeros(m, np.float32)
    L.msflref_knn_log_copy(_ptr(k_of, C.c_int), _ptr(idx, C.c_int), _ptr(d2, C.c_float))
    return k_of, idx, d2


def transform_point(pose, xyz):
    """TransformPoint (rigid_transform.h:132-138) on an (n, 3) float array."""
    xyz = _f32(xyz)[:, :3].copy()
    out = np.zeros_like(xyz)
    x = _pose(pose)
    f = ref_lib().msflref_transform_point
    for i in range(xyz.shape[0]):
        f(_ptr(x, C.c_double), _ptr(xyz[i], C.c_float), _ptr(out[i], C.c_float))
    return out


def get_delta_qp(sum_dt, delta_q, delta_p, dt):
    t, q, p = _d(sum_dt), _d(delta_q).reshape(-1, 4), _d(delta_p).reshape(-1, 3)
    dq, dp = np.zeros(4), np.zeros(3)
    ref_lib().msflref_get_delta_qp(_ptr(t, C.c_double), _ptr(q, C.c_double), _ptr(p, C.c_double), C.c_int(t.shape[0]),
                                   C.c_double(dt), _ptr(dq, C.c_double), _ptr(dp, C.c_double))
    return dq, dp


def scan2map(map_corner, map_surf, scan_corner, scan_surf, pose):
    """MappingScanMatcher::MatchScan2Map, is_initialized == false.  Returns (ok, pose)."""
    mc, ms, sc, ss = (_f32(a, 4) for a in (map_corner, map_surf, scan_corner, scan_surf))
    x = _pose(pose)
    ok = ref_lib().msflref_scan2map(_ptr(mc, C.c_float), C.c_int(mc.shape[0]), _ptr(ms, C.c_float), C.c_int(ms.shape[0]),
                                    _ptr(sc, C.c_float), C.c_int(sc.shape[0]), _ptr(ss, C.c_float), C.c_int(ss.shape[0]),
                                    _ptr(x, C.c_double))
    return bool(ok), x


def scan2map_deskew(map_corner, map_surf, scan_corner, scan_surf, sum_dt, delta_q, delta_p, velocity, gravity, pose):
    """MatchScan2Map, is_initialized == true, entered after the IMU-only predict.  Returns (ok, pose, velocity)."""
    mc, ms, sc, ss = (_f32(a, 4) for a in (map_corner, map_surf, scan_corner, scan_surf))
    t, q, p = _d(sum_dt), _d(delta_q).reshape(-1, 4), _d(delta_p).reshape(-1, 3)
    V, G = _d(velocity), _d(gravity)
    x, v = _pose(pose), np.zeros(3)
    ok = ref_lib().msflref_scan2map_deskew(_ptr(mc, C.c_float), C.c_int(mc.shape[0]), _ptr(ms, C.c_float), C.c_int(ms.shape[0]),
                                           _ptr(sc, C.c_float), C.c_int(sc.shape[0]), _ptr(ss, C.c_float), C.c_int(ss.shape[0]),
                                           _ptr(t, C.c_double), _ptr(q, C.c_double), _ptr(p, C.c_double), C.c_int(t.shape[0]),
                                           _ptr(V, C.c_double), _ptr(G, C.c_double), _ptr(x, C.c_double), _ptr(v, C.c_double))
    return bool(ok), x, v


def scan2scan(last_corner, last_corner_ring, last_surf, last_surf_ring, curr_sharp, curr_flat, pose):
    """OdometryScanMatcher::MatchScan2Scan.  Returns (ok, pose): ok False = fewer than 10 correspondences."""
    lc, ls, cs, cf = (_f32(a, 4) for a in (last_corner, last_surf, curr_sharp, curr_flat))
    lcr = np.ascontiguousarray(last_corner_ring, dtype=np.uint16)
    lsr = np.ascontiguousarray(last_surf_ring, dtype=np.uint16)
    x = _pose(pose)
    ok = ref_lib().msflref_scan2scan(_ptr(lc, C.c_float), _ptr(lcr, C.c_uint16), C.c_int(lc.shape[0]), _ptr(ls, C.c_float),
                                     _ptr(lsr, C.c_uint16), C.c_int(ls.shape[0]), _ptr(cs, C.c_float), C.c_int(cs.shape[0]),
                                     _ptr(cf, C.c_float), C.c_int(cf.shape[0]), _ptr(x, C.c_double))
    return bool(ok), x


def extract_features(xyzi, ring, T_ext=None, min_range=0.3):
    """RealHandleLaserCloudMessage (msf_loam_node.cc:160-378).  Returns the registered full cloud, its rings and the four
    feature clouds in the reference's push order."""
    pts = _f32(xyzi, 4)
    rg = np.ascontiguousarray(ring, dtype=np.uint16)
    n = pts.shape[0]
    T = _pose(T_ext if T_ext is not None else [0, 0, 0, 0, 0, 0, 1])
    names = ("full", "sharp", "less_sharp", "flat", "less_flat")
    bufs = {k: np.zeros((n, 4), np.float32) for k in names}
    cnt = {k: C.c_int() for k in names}
    full_ring = np.zeros(n, np.uint16)
    rc = ref_lib().msflref_extract_features(
        _ptr(pts, C.c_float), _ptr(rg, C.c_uint16), C.c_int(n), _ptr(T, C.c_double), C.c_double(min_range),
        _ptr(bufs["full"], C.c_float), _ptr(full_ring, C.c_uint16), C.byref(cnt["full"]),
        _ptr(bufs["sharp"], C.c_float), C.byref(cnt["sharp"]), _ptr(bufs["less_sharp"], C.c_float), C.byref(cnt["less_sharp"]),
        _ptr(bufs["flat"], C.c_float), C.byref(cnt["flat"]), _ptr(bufs["less_flat"], C.c_float), C.byref(cnt["less_flat"]))
    assert rc == 1
    out = {k: bufs[k][:cnt[k].value] for k in names}
    out["ring"] = full_ring[:cnt["full"].value]
    return out


def scan2map_batch(map_corner, map_surf, scan_corner, corner_off, scan_surf, surf_off, poses, n_threads=1, fixed_attempts=0):
    """B MatchScan2Map calls (LiDAR-only) on n_threads host threads; fixed_attempts > 0 = bench.py's fixed schedule.
    Returns poses (B, 7)."""
    mc, ms, sc, ss = (_f32(a, 4) for a in (map_corner, map_surf, scan_corner, scan_surf))
    co = np.ascontiguousarray(corner_off, dtype=np.int32)
    so = np.ascontiguousarray(surf_off, dtype=np.int32)
    B = co.shape[0] - 1
    x = np.ascontiguousarray(poses, dtype=np.float64).reshape(B, 7).copy()
    ref_lib().msflref_scan2map_batch(_ptr(mc, C.c_float), C.c_int(mc.shape[0]), _ptr(ms, C.c_float), C.c_int(ms.shape[0]),
                                     C.c_int(B), _ptr(sc, C.c_float), _ptr(co, C.c_int), _ptr(ss, C.c_float), _ptr(so, C.c_int),
                                     _ptr(x, C.c_double), C.c_int(n_threads), C.c_int(fixed_attempts))
    return x


class Map:
    """The reference's HybridGrid (slam/map/hybrid_grid.cc) with the caller's pcl::VoxelGrid of the given leaf size."""

    def __init__(self, resolution=3.0, leaf=0.2):
        L = ref_lib()
        L.msflref_map_create.restype = C.c_void_p
        self.h = C.c_void_p(L.msflref_map_create(C.c_float(resolution), C.c_float(leaf)))

    def __del__(self):
        if getattr(self, "h", None):
            ref_lib().msflref_map_free(self.h)
            self.h = None

    def insert(self, scan_world_xyzi):
        pts = _f32(scan_world_xyzi, 4)
        ref_lib().msflref_map_insert(self.h, _ptr(pts, C.c_float), C.c_int(pts.shape[0]))

    def surround(self, scan_xyzi, pose, cap=2_000_000):
        """GetSurroundedCloud: (n, 4) float array, cells in the reference's (heap-address) order."""
        pts = _f32(scan_xyzi, 4)
        out = np.zeros((cap, 4), np.float32)
        x = _pose(pose)
        n = ref_lib().msflref_map_surround(self.h, _ptr(pts, C.c_float), C.c_int(pts.shape[0]), _ptr(x, C.c_double),
                                           _ptr(out, C.c_float), C.c_int(cap))
        assert n <= cap
        return out[:n].copy()
