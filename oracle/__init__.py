"""CPU oracle (TEST INFRASTRUCTURE ONLY -- pinned to the reference's own compiled code in oracle/_ref, third-party numerics
restated; see msfl_oracle.h and ref_shim.cc).

ctypes bindings over ``libmsfl_oracle.so`` (plain-C restatement of the reference hot path).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  ``msf_loam_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsfl_oracle.so")
MAX_ATTEMPTS = 64
CORR_STRIDE = 10


class Params(C.Structure):
    _fields_ = [
        ("min_range", C.c_double), ("scan_period", C.c_double), ("curvature_thresh", C.c_double),
        ("neighbor_gap_sq", C.c_double), ("n_sectors", C.c_int), ("n_sharp", C.c_int),
        ("n_less_sharp", C.c_int), ("n_flat", C.c_int),
        ("dist_sq_thresh", C.c_double), ("nearby_scan", C.c_double), ("min_correspondences", C.c_int),
        ("knn_max_sq", C.c_double), ("line_eig_ratio", C.c_double), ("line_half_len", C.c_double),
        ("plane_tol", C.c_double),
        ("num_outer", C.c_int), ("max_num_iterations", C.c_int), ("huber_a", C.c_double),
        ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
        ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
        ("max_lm_diagonal", C.c_double), ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("max_consecutive_invalid_steps", C.c_int), ("early_exit", C.c_int),
    ]


class LmIter(C.Structure):
    _fields_ = [("cost", C.c_double), ("cost_candidate", C.c_double), ("model_change", C.c_double),
                ("rho", C.c_double), ("radius", C.c_double), ("valid", C.c_int), ("accepted", C.c_int)]


class LmLog(C.Structure):
    _fields_ = [("n_attempts", C.c_int), ("termination", C.c_int), ("initial_cost", C.c_double),
                ("final_cost", C.c_double), ("it", LmIter * MAX_ATTEMPTS)]

    def as_dict(self):
        return {
            "n_attempts": self.n_attempts, "termination": self.termination,
            "initial_cost": self.initial_cost, "final_cost": self.final_cost,
            "iters": [
                {k: getattr(self.it[i], k) for k, _ in LmIter._fields_} for i in range(self.n_attempts)
            ],
        }


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc); returns the .so path."""
    src = os.path.join(_HERE, "msfl_oracle.c")
    hdr = os.path.join(_HERE, "msfl_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libmsfl_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libmsfl_ref.so")
REFERENCE_ROOT = os.environ.get("MSFL_REFERENCE_ROOT", "/root/reference")


def build_ref(force: bool = False):
    """Compile the REFERENCE's own hot-path sources (oracle/Makefile, target ``ref``: msf_loam_node.cc, lidar_factor.cc,
    odometry_scan_matcher.cc, mapping_scan_matcher.cc, scan_matcher.cc, pose_local_parameterization.cc,
    scan_undistortion.cc, unmodified, against the stand-in headers of oracle/ref_stubs/) from the checkout at
    REFERENCE_ROOT into oracle/_ref/.  Returns the .so path, or None when there is neither a
    checkout nor a prebuilt library (the GPU box only ever uses the prebuilt file)."""
    srcs = [os.path.join(REFERENCE_ROOT, "src/slam/local/scan_matching", f) for f in
            ("lidar_factor.cc", "odometry_scan_matcher.cc", "mapping_scan_matcher.cc", "scan_matcher.cc")]
    srcs += [os.path.join(REFERENCE_ROOT, "src/slam/imu_fusion", f) for f in
             ("pose_local_parameterization.cc", "scan_undistortion.cc")]
    srcs.append(os.path.join(REFERENCE_ROOT, "src/msf_loam_node.cc"))
    srcs.append(os.path.join(REFERENCE_ROOT, "src/slam/map/hybrid_grid.cc"))
    if not all(os.path.exists(f) for f in srcs):
        return _REF_LIB_PATH if os.path.exists(_REF_LIB_PATH) else None
    cmd = ["make", "-C", _HERE, "REF=" + REFERENCE_ROOT, "ref"] + (["-B"] if force else [])
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return _REF_LIB_PATH


_ref_lib = None


def ref_lib():
    """ctypes handle of oracle/_ref/libmsfl_ref.so (the reference's compiled scan-matching code), or None."""
    global _ref_lib
    if _ref_lib is None:
        path = build_ref()
        if path is None:
            return None
        _ref_lib = C.CDLL(path)
    return _ref_lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.msflo_kdtree_build.restype = C.c_void_p
    return _lib


def default_params(**over) -> Params:
    p = Params()
    lib().msflo_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def _f32(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _pose(p):
    return np.ascontiguousarray(p, dtype=np.float64).reshape(7).copy()


def pose_plus(x, delta):
    out = np.zeros(7)
    lib().msflo_pose_plus(_ptr(_pose(x), C.c_double),
                          _ptr(np.ascontiguousarray(delta, dtype=np.float64), C.c_double),
                          _ptr(out, C.c_double))
    return out


def transform_points_f(pose, xyz):
    xyz = _f32(xyz)
    out = np.zeros((xyz.shape[0], 3), np.float32)
    pz = _pose(pose)
    tmp_in = np.zeros(3, np.float32)
    tmp_out = np.zeros(3, np.float32)
    f = lib().msflo_transform_point_f
    for i in range(xyz.shape[0]):
        tmp_in[:] = xyz[i, :3]
        f(_ptr(pz, C.c_double), _ptr(tmp_in, C.c_float), _ptr(tmp_out, C.c_float))
        out[i] = tmp_out
    return out


def edge_factor(pose, p, a, n):
    r = np.zeros(3)
    J = np.zeros(21)
    args = [np.ascontiguousarray(v, dtype=np.float64) for v in (p, a, n)]
    lib().msflo_edge_factor(_ptr(_pose(pose), C.c_double), *[_ptr(v, C.c_double) for v in args],
                            _ptr(r, C.c_double), _ptr(J, C.c_double))
    return r, J.reshape(3, 7)


def plane_factor(pose, p, c, n):
    r = np.zeros(1)
    J = np.zeros(7)
    args = [np.ascontiguousarray(v, dtype=np.float64) for v in (p, c, n)]
    lib().msflo_plane_factor(_ptr(_pose(pose), C.c_double), *[_ptr(v, C.c_double) for v in args],
                             _ptr(r, C.c_double), _ptr(J, C.c_double))
    return r, J.reshape(1, 7)


def accumulate(params, corr, pose):
    corr = np.ascontiguousarray(corr, dtype=np.float64).reshape(-1, CORR_STRIDE)
    cost = C.c_double(0)
    H = np.zeros(36)
    g = np.zeros(6)
    lib().msflo_accumulate(C.byref(params), _ptr(corr, C.c_double), C.c_int(corr.shape[0]),
                           _ptr(_pose(pose), C.c_double), C.byref(cost), _ptr(H, C.c_double),
                           _ptr(g, C.c_double))
    return cost.value, H.reshape(6, 6), g


def lm_solve(params, corr, pose):
    corr = np.ascontiguousarray(corr, dtype=np.float64).reshape(-1, CORR_STRIDE)
    x = _pose(pose)
    log = LmLog()
    lib().msflo_lm_solve(C.byref(params), _ptr(corr, C.c_double), C.c_int(corr.shape[0]),
                         _ptr(x, C.c_double), C.byref(log))
    return x, log.as_dict()


def knn(points_xyzi, queries_xyz, k, brute=False):
    pts = _f32(points_xyzi, 4)
    q = _f32(queries_xyz)[:, :3].copy()
    idx = np.full((q.shape[0], k), -1, np.int32)
    d2 = np.zeros((q.shape[0], k), np.float32)
    fn = lib().msflo_knn_brute if brute else lib().msflo_knn_batch
    fn(_ptr(pts, C.c_float), C.c_int(pts.shape[0]), _ptr(q, C.c_float), C.c_int(q.shape[0]),
       C.c_int(k), _ptr(idx, C.c_int), _ptr(d2, C.c_float))
    return idx, d2


def sym_eig3(A):
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(9)
    ev = np.zeros(3)
    V = np.zeros(9)
    lib().msflo_sym_eig3(_ptr(A, C.c_double), _ptr(ev, C.c_double), _ptr(V, C.c_double))
    return ev, V.reshape(3, 3)


def lstsq_5x3(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64).reshape(15)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(5)
    x = np.zeros(3)
    lib().msflo_lstsq_5x3(_ptr(A, C.c_double), _ptr(b, C.c_double), _ptr(x, C.c_double))
    return x


def associate_map(params, map_corner, map_surf, scan_corner, scan_surf, pose):
    """Returns (corr [n,10], n_edge, n_plane, knn_idx [(nc+ns),5])."""
    mc, ms, sc, ss = (_f32(a, 4) for a in (map_corner, map_surf, scan_corner, scan_surf))
    L = lib()
    tc = C.c_void_p(L.msflo_kdtree_build(_ptr(mc, C.c_float), C.c_int(mc.shape[0])))
    ts = C.c_void_p(L.msflo_kdtree_build(_ptr(ms, C.c_float), C.c_int(ms.shape[0])))
    nq = sc.shape[0] + ss.shape[0]
    corr = np.zeros((nq + 1, CORR_STRIDE))
    ne, npl = C.c_int(0), C.c_int(0)
    kidx = np.full((nq, 5), -1, np.int32)
    L.msflo_associate_map(C.byref(params), tc, _ptr(mc, C.c_float), ts, _ptr(ms, C.c_float),
                          _ptr(sc, C.c_float), C.c_int(sc.shape[0]), _ptr(ss, C.c_float),
                          C.c_int(ss.shape[0]), _ptr(_pose(pose), C.c_double), _ptr(corr, C.c_double),
                          C.byref(ne), C.byref(npl), _ptr(kidx, C.c_int))
    L.msflo_kdtree_free(tc)
    L.msflo_kdtree_free(ts)
    return corr[: ne.value + npl.value].copy(), ne.value, npl.value, kidx


def scan2map(params, map_corner, map_surf, scan_corner, scan_surf, pose):
    """MappingScanMatcher::MatchScan2Map (LiDAR-only).  Returns (pose, logs, counts)."""
    mc, ms, sc, ss = (_f32(a, 4) for a in (map_corner, map_surf, scan_corner, scan_surf))
    x = _pose(pose)
    logs = (LmLog * params.num_outer)()
    counts = np.zeros(2 * params.num_outer, np.int32)
    lib().msflo_scan2map(C.byref(params), _ptr(mc, C.c_float), C.c_int(mc.shape[0]),
                         _ptr(ms, C.c_float), C.c_int(ms.shape[0]),
                         _ptr(sc, C.c_float), C.c_int(sc.shape[0]), _ptr(ss, C.c_float),
                         C.c_int(ss.shape[0]), _ptr(x, C.c_double), logs, _ptr(counts, C.c_int))
    return x, [l.as_dict() for l in logs], counts.reshape(-1, 2)


class Deskew(C.Structure):
    _fields_ = [("sum_dt", C.POINTER(C.c_double)), ("delta_q", C.POINTER(C.c_double)),
                ("delta_p", C.POINTER(C.c_double)), ("n", C.c_int), ("velocity", C.c_double * 3),
                ("gravity", C.c_double * 3)]


def scan2map_deskew(params, map_corner, map_surf, scan_corner, scan_surf, sum_dt, delta_q, delta_p, velocity,
                    gravity, pose):
    """IMU-initialised branch of MatchScan2Map (LiDAR part).  Returns (rc, pose, logs, counts, knn_idx)."""
    mc, ms, sc, ss = (_f32(a, 4) for a in (map_corner, map_surf, scan_corner, scan_surf))
    t = np.ascontiguousarray(sum_dt, dtype=np.float64)
    q = np.ascontiguousarray(delta_q, dtype=np.float64).reshape(-1, 4)
    p = np.ascontiguousarray(delta_p, dtype=np.float64).reshape(-1, 3)
    dk = Deskew(_ptr(t, C.c_double), _ptr(q, C.c_double), _ptr(p, C.c_double), t.shape[0],
                (C.c_double * 3)(*velocity), (C.c_double * 3)(*gravity))
    x = _pose(pose)
    logs = (LmLog * params.num_outer)()
    counts = np.zeros(2 * params.num_outer, np.int32)
    kidx = np.full((sc.shape[0] + ss.shape[0], 5), -1, np.int32)
    rc = lib().msflo_scan2map_deskew(C.byref(params), _ptr(mc, C.c_float), C.c_int(mc.shape[0]),
                                     _ptr(ms, C.c_float), C.c_int(ms.shape[0]), _ptr(sc, C.c_float),
                                     C.c_int(sc.shape[0]), _ptr(ss, C.c_float), C.c_int(ss.shape[0]),
                                     C.byref(dk), _ptr(x, C.c_double), logs, _ptr(counts, C.c_int),
                                     _ptr(kidx, C.c_int))
    return rc, x, [l.as_dict() for l in logs], counts.reshape(-1, 2), kidx


def scan2map_batch(params, map_corner, map_surf, scan_corner, corner_off, scan_surf, surf_off, poses,
                   n_threads=1):
    mc, ms, sc, ss = (_f32(a, 4) for a in (map_corner, map_surf, scan_corner, scan_surf))
    co = np.ascontiguousarray(corner_off, dtype=np.int32)
    so = np.ascontiguousarray(surf_off, dtype=np.int32)
    B = co.shape[0] - 1
    x = np.ascontiguousarray(poses, dtype=np.float64).reshape(B, 7).copy()
    lib().msflo_scan2map_batch(C.byref(params), _ptr(mc, C.c_float), C.c_int(mc.shape[0]),
                               _ptr(ms, C.c_float), C.c_int(ms.shape[0]), C.c_int(B),
                               _ptr(sc, C.c_float), _ptr(co, C.c_int), _ptr(ss, C.c_float),
                               _ptr(so, C.c_int), _ptr(x, C.c_double), C.c_int(n_threads))
    return x


def scan2scan(params, last_corner, last_corner_ring, last_surf, last_surf_ring, curr_sharp, curr_flat,
              pose):
    """OdometryScanMatcher::MatchScan2Scan.  Returns (status, pose, logs, counts, assoc)."""
    lc, ls, cs, cf = (_f32(a, 4) for a in (last_corner, last_surf, curr_sharp, curr_flat))
    lcr = np.ascontiguousarray(last_corner_ring, dtype=np.uint16)
    lsr = np.ascontiguousarray(last_surf_ring, dtype=np.uint16)
    x = _pose(pose)
    logs = (LmLog * params.num_outer)()
    counts = np.zeros(2 * params.num_outer, np.int32)
    assoc = np.full(2 * cs.shape[0] + 3 * cf.shape[0] + 1, -1, np.int32)
    rc = lib().msflo_scan2scan(C.byref(params), _ptr(lc, C.c_float), _ptr(lcr, C.c_uint16),
                               C.c_int(lc.shape[0]), _ptr(ls, C.c_float), _ptr(lsr, C.c_uint16),
                               C.c_int(ls.shape[0]), _ptr(cs, C.c_float), C.c_int(cs.shape[0]),
                               _ptr(cf, C.c_float), C.c_int(cf.shape[0]), _ptr(x, C.c_double), logs,
                               _ptr(counts, C.c_int), _ptr(assoc, C.c_int))
    return rc, x, [l.as_dict() for l in logs], counts.reshape(-1, 2), assoc[:-1]


def extract_features(params, xyzi, ring, T_ext=None):
    """RealHandleLaserCloudMessage feature block.  Returns dict of arrays."""
    pts = _f32(xyzi, 4)
    rg = np.ascontiguousarray(ring, dtype=np.uint16)
    n = pts.shape[0]
    full = np.zeros((n, 4), np.float32)
    fring = np.zeros(n, np.uint16)
    curv = np.zeros(n, np.float32)
    label = np.zeros(n, np.int32)
    idx = [np.zeros(n, np.int32) for _ in range(4)]
    cnt = [C.c_int(0) for _ in range(5)]
    T = _pose(T_ext) if T_ext is not None else None
    rc = lib().msflo_extract_features(
        C.byref(params), _ptr(pts, C.c_float), _ptr(rg, C.c_uint16), C.c_int(n),
        _ptr(T, C.c_double) if T is not None else None,
        _ptr(full, C.c_float), _ptr(fring, C.c_uint16), C.byref(cnt[0]),
        _ptr(curv, C.c_float), _ptr(label, C.c_int),
        _ptr(idx[0], C.c_int), C.byref(cnt[1]), _ptr(idx[1], C.c_int), C.byref(cnt[2]),
        _ptr(idx[2], C.c_int), C.byref(cnt[3]), _ptr(idx[3], C.c_int), C.byref(cnt[4]))
    nf = cnt[0].value
    return {
        "status": rc, "full": full[:nf], "ring": fring[:nf], "curvature": curv[:nf], "label": label[:nf],
        "idx_sharp": idx[0][: cnt[1].value], "idx_less_sharp": idx[1][: cnt[2].value],
        "idx_flat": idx[2][: cnt[3].value], "idx_less_flat": idx[3][: cnt[4].value],
    }


def voxel_grid(xyzi, leaf):
    pts = _f32(xyzi, 4)
    out = np.zeros_like(pts)
    n = lib().msflo_voxel_grid(_ptr(pts, C.c_float), C.c_int(pts.shape[0]), C.c_float(leaf),
                               _ptr(out, C.c_float))
    return out[:n].copy()


class Stgm:
    """HybridGrid restatement (hybrid_grid.cc:403-521); output order = ascending cell key."""

    def __init__(self, resolution=3.0, leaf=0.4):
        L = lib()
        L.msflo_stgm_create.restype = C.c_void_p
        self.h = C.c_void_p(L.msflo_stgm_create(C.c_float(resolution), C.c_float(leaf)))

    def __del__(self):
        try:
            lib().msflo_stgm_free(self.h)
        except Exception:
            pass

    def insert(self, scan_world_xyzi):
        pts = _f32(scan_world_xyzi, 4)
        lib().msflo_stgm_insert(self.h, _ptr(pts, C.c_float), C.c_int(pts.shape[0]))

    def size(self):
        nc = C.c_int(0)
        n = lib().msflo_stgm_size(self.h, C.byref(nc))
        return n, nc.value

    def surround(self, scan_xyzi, pose):
        pts = _f32(scan_xyzi, 4)
        out = np.zeros((max(self.size()[0], 1), 4), np.float32)
        n = lib().msflo_stgm_surround(self.h, _ptr(pts, C.c_float), C.c_int(pts.shape[0]),
                                      _ptr(_pose(pose), C.c_double), _ptr(out, C.c_float))
        return out[:n].copy()

    def dump(self):
        out = np.zeros((max(self.size()[0], 1), 4), np.float32)
        n = lib().msflo_stgm_dump(self.h, _ptr(out, C.c_float))
        return out[:n].copy()
