"""Host-side mirror of the reference's scan-matching interface over the C ABI (include/msfl.h).

``MappingScanMatcher.MatchScan2Map`` / ``OdometryScanMatcher.MatchScan2Scan`` /
``ScanRegistration.extract`` keep the reference's names, argument meaning and error behaviour
(mapping_scan_matcher.h:14-21, odometry_scan_matcher.h:10-12, msf_loam_node.cc:160-371) so the
parity tests read like tests of the reference.  All compute happens in libmsfl.so on the GPU;
this module only marshals numpy arrays into ``msfl_cloud`` views.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Cloud, Deskew, Features, MsflError, Params, Stats, NO_FIELD, MSFL_OK, MSFL_TOO_FEW

# memory layouts of the reference's point types (common.h:44-62; pcl::PointXYZI)
POINT_XYZI = np.dtype({"names": ["x", "y", "z", "intensity"], "formats": ["f4"] * 4,
                       "offsets": [0, 4, 8, 16], "itemsize": 32})
POINT_XYZIRT = np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                         "formats": ["f4", "f4", "f4", "f4", "u2", "f4"],
                         "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32})


def default_params(**over) -> Params:
    p = Params()
    _lib.load_library().msfl_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def to_pcl(xyzi, ring=None):
    """numpy (n,4) float32 [+ ring] -> structured array with the reference's PCL point layout."""
    xyzi = np.asarray(xyzi, dtype=np.float32).reshape(-1, 4)
    out = np.zeros(xyzi.shape[0], dtype=POINT_XYZIRT if ring is not None else POINT_XYZI)
    out["x"], out["y"], out["z"], out["intensity"] = xyzi[:, 0], xyzi[:, 1], xyzi[:, 2], xyzi[:, 3]
    if ring is not None:
        out["ring"] = np.asarray(ring, dtype=np.uint16)
    return out


class _View:
    """Keeps the numpy buffer alive next to the msfl_cloud that points into it."""

    def __init__(self, arr, ring=None):
        if isinstance(arr, _View):
            self.buf, self.cloud, self.aux = arr.buf, arr.cloud, arr.aux
            return
        self.aux = None
        a = np.asarray(arr)
        if a.dtype.names:  # structured (PCL layout)
            a = np.ascontiguousarray(a)
            offs = {n: a.dtype.fields[n][1] for n in a.dtype.names}
            self.buf = a
            self.cloud = Cloud(a.ctypes.data, a.shape[0], a.dtype.itemsize, offs["x"],
                               offs.get("intensity", NO_FIELD), offs.get("ring", NO_FIELD))
        elif ring is not None:
            self.buf = to_pcl(a, ring)
            offs = {n: self.buf.dtype.fields[n][1] for n in self.buf.dtype.names}
            self.cloud = Cloud(self.buf.ctypes.data, self.buf.shape[0], self.buf.dtype.itemsize,
                               offs["x"], offs["intensity"], offs["ring"])
        elif a.ndim == 2 and a.shape[1] == 3:  # xyz only (12 B points): enough for the LiDAR-only matcher's queries
            a = np.ascontiguousarray(a, dtype=np.float32)
            self.buf = a
            self.cloud = Cloud(a.ctypes.data, a.shape[0], 12, 0, NO_FIELD, NO_FIELD)
        else:
            a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4)
            self.buf = a
            self.cloud = Cloud(a.ctypes.data, a.shape[0], 16, 0, 12, NO_FIELD)

    @property
    def n(self):
        return int(self.cloud.n)


def _pose(p):
    return np.ascontiguousarray(p, dtype=np.float64).reshape(-1).copy()


class Engine:
    """One msfl_engine (one CUDA stream + device buffers); one per matcher thread."""

    def __init__(self, params: Params | None = None, device: int = 0, stream: int | None = None):
        self.lib = _lib.load_library()
        self.params = params if params is not None else default_params()
        h = C.c_void_p()
        rc = self.lib.msfl_create_on_stream(C.byref(self.params), C.c_int(device),
                                            C.c_void_p(stream) if stream else None, C.byref(h))
        self._check(rc)
        self.h = h
        self.device = device

    def _check(self, rc):
        if rc < 0:
            raise MsflError(rc, self.lib.msfl_last_error().decode())
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.msfl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.msfl_sync(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.msfl_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.msfl_launch_count(self.h))

    def set_profiling(self, on: bool):
        self._check(self.lib.msfl_set_profiling(self.h, C.c_int(1 if on else 0)))

    def get_profile(self):
        """(ms[stage], count[stage]) since the last call; stages: 0 association, 1 solver, 2 sort."""
        ms = (C.c_double * _lib.N_STAGES)()
        cnt = (C.c_int32 * _lib.N_STAGES)()
        self._check(self.lib.msfl_get_profile(self.h, ms, cnt))
        return list(ms), list(cnt)

    # ---------------------------------------------------------------- submap
    def set_submap(self, map_corner, map_surf):
        vc, vs = _View(map_corner), _View(map_surf)
        self._check(self.lib.msfl_set_submap(self.h, C.byref(vc.cloud), C.byref(vs.cloud)))

    def set_submap_device(self, d_corner_ptr: int, n_corner: int, d_surf_ptr: int, n_surf: int):
        self._check(self.lib.msfl_set_submap_device(self.h, C.c_void_p(d_corner_ptr), C.c_size_t(n_corner),
                                                    C.c_void_p(d_surf_ptr), C.c_size_t(n_surf)))

    def get_submap_device(self):
        pc, ps = C.c_void_p(), C.c_void_p()
        nc, ns = C.c_size_t(), C.c_size_t()
        self._check(self.lib.msfl_get_submap_device(self.h, C.byref(pc), C.byref(nc), C.byref(ps), C.byref(ns)))
        return pc.value, nc.value, ps.value, ns.value

    # ---------------------------------------------------------------- multi-GPU: submap broadcast (NCCL, C ABI)
    def nccl_comm_init(self, unique_id: bytes, nranks: int, rank: int) -> int:
        """msfl_nccl_comm_init: returns the ncclComm_t as an int handle (every rank passes rank 0's id)."""
        assert len(unique_id) == _lib.NCCL_UNIQUE_ID_BYTES
        buf = (C.c_ubyte * _lib.NCCL_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        comm = C.c_void_p()
        self._check(self.lib.msfl_nccl_comm_init(self.h, buf, C.c_int(nranks), C.c_int(rank), C.byref(comm)))
        return comm.value

    def nccl_comm_destroy(self, comm: int):
        self._check(self.lib.msfl_nccl_comm_destroy(C.c_void_p(comm)))

    def bcast_submap(self, comm: int, root: int = 0):
        """msfl_bcast_submap: collective; afterwards every rank holds root's submap AND its cell index."""
        self._check(self.lib.msfl_bcast_submap(self.h, C.c_void_p(comm), C.c_int(root)))

    # ---------------------------------------------------------------- scan-to-map
    def scan2map(self, scan_corner, scan_surf, pose, want_stats=True):
        vc, vs = _View(scan_corner), _View(scan_surf)
        x = _pose(pose)
        st = Stats() if want_stats else None
        rc = self._check(self.lib.msfl_scan2map(self.h, C.byref(vc.cloud), C.byref(vs.cloud),
                                                x.ctypes.data_as(C.POINTER(C.c_double)),
                                                C.byref(st) if st is not None else None))
        return rc, x, (st.as_dict() if st is not None else None)

    def scan2map_batch(self, scan_corners, scan_surfs, poses, want_stats=False):
        B = len(scan_corners)
        vcs = [_View(a) for a in scan_corners]
        vss = [_View(a) for a in scan_surfs]
        carr = (Cloud * B)(*[v.cloud for v in vcs])
        sarr = (Cloud * B)(*[v.cloud for v in vss])
        x = np.ascontiguousarray(poses, dtype=np.float64).reshape(B, 7).copy()
        st = (Stats * B)() if want_stats else None
        rc = self._check(self.lib.msfl_scan2map_batch(self.h, C.c_int(B), carr, sarr,
                                                      x.ctypes.data_as(C.POINTER(C.c_double)), st))
        return rc, x, ([s.as_dict() for s in st] if st is not None else None)

    def scan2map_deskew(self, scan_corner, scan_surf, sum_dt, delta_q, delta_p, velocity, gravity, pose,
                        want_stats=True):
        """IMU-initialised branch: preintegration buffers (sum_dt [n], delta_q [n,4] xyzw, delta_p [n,3])."""
        vc, vs = _View(scan_corner), _View(scan_surf)
        t = np.ascontiguousarray(sum_dt, dtype=np.float64)
        q = np.ascontiguousarray(delta_q, dtype=np.float64).reshape(-1, 4)
        p = np.ascontiguousarray(delta_p, dtype=np.float64).reshape(-1, 3)
        dk = Deskew(t.ctypes.data_as(C.POINTER(C.c_double)), q.ctypes.data_as(C.POINTER(C.c_double)),
                    p.ctypes.data_as(C.POINTER(C.c_double)), t.shape[0], 0,
                    (C.c_double * 3)(*velocity), (C.c_double * 3)(*gravity))
        x = _pose(pose)
        st = Stats() if want_stats else None
        rc = self._check(self.lib.msfl_scan2map_deskew(self.h, C.byref(vc.cloud), C.byref(vs.cloud), C.byref(dk),
                                                       x.ctypes.data_as(C.POINTER(C.c_double)),
                                                       C.byref(st) if st is not None else None))
        return rc, x, (st.as_dict() if st is not None else None)

    def prepare_deskew_batch(self, scan_corners, scan_surfs, tables):
        """Builds the msfl_cloud / msfl_deskew tables of a Deskew-branch batch once (tables[b] = (sum_dt, delta_q,
        delta_p, velocity, gravity) of scan b), so a timed loop only pays the C call."""
        B = len(scan_corners)
        vcs, vss = [_View(a) for a in scan_corners], [_View(a) for a in scan_surfs]
        keep, dks = [], (Deskew * B)()
        for b, (sum_dt, delta_q, delta_p, velocity, gravity) in enumerate(tables):
            t = np.ascontiguousarray(sum_dt, dtype=np.float64)
            q = np.ascontiguousarray(delta_q, dtype=np.float64).reshape(-1, 4)
            p = np.ascontiguousarray(delta_p, dtype=np.float64).reshape(-1, 3)
            keep.append((t, q, p))
            dks[b] = Deskew(t.ctypes.data_as(C.POINTER(C.c_double)), q.ctypes.data_as(C.POINTER(C.c_double)),
                            p.ctypes.data_as(C.POINTER(C.c_double)), t.shape[0], 0,
                            (C.c_double * 3)(*velocity), (C.c_double * 3)(*gravity))
        return {"B": B, "views": (vcs, vss), "tables": keep, "carr": (Cloud * B)(*[v.cloud for v in vcs]),
                "sarr": (Cloud * B)(*[v.cloud for v in vss]), "dks": dks}

    def scan2map_deskew_prepared(self, batch, poses_inout: np.ndarray, stats=None):
        """msfl_scan2map_deskew_batch on a prepared batch; poses_inout (B, 7) float64 is updated in place."""
        assert poses_inout.dtype == np.float64 and poses_inout.flags.c_contiguous
        return self._check(self.lib.msfl_scan2map_deskew_batch(
            self.h, C.c_int(batch["B"]), batch["carr"], batch["sarr"], batch["dks"],
            poses_inout.ctypes.data_as(C.POINTER(C.c_double)), stats))

    def scan2map_deskew_batch(self, scan_corners, scan_surfs, tables, poses, want_stats=True):
        """msfl_scan2map_deskew_batch: B scans of a replayed log through the IMU-initialised branch in one call.
        tables[b] = (sum_dt, delta_q, delta_p, velocity, gravity) of scan b; poses (B, 7) = the poses after each scan's
        IMU-only predict.  Returns (rc, poses (B, 7), list of stats dicts or None)."""
        batch = self.prepare_deskew_batch(scan_corners, scan_surfs, tables)
        x = np.ascontiguousarray(poses, dtype=np.float64).reshape(batch["B"], 7).copy()
        st = (Stats * batch["B"])() if want_stats else None
        rc = self.scan2map_deskew_prepared(batch, x, st)
        return rc, x, ([s.as_dict() for s in st] if st is not None else None)

    def prepare_batch(self, scan_corners, scan_surfs):
        """Builds the msfl_cloud tables once so a timed loop only pays the C call."""
        B = len(scan_corners)
        vcs = [_View(a) for a in scan_corners]
        vss = [_View(a) for a in scan_surfs]
        return {"B": B, "views": (vcs, vss), "carr": (Cloud * B)(*[v.cloud for v in vcs]),
                "sarr": (Cloud * B)(*[v.cloud for v in vss])}

    def scan2map_prepared(self, batch, poses_inout: np.ndarray):
        """msfl_scan2map_batch on a prepared batch; poses_inout (B,7) float64 is updated in place."""
        assert poses_inout.dtype == np.float64 and poses_inout.flags.c_contiguous
        return self._check(self.lib.msfl_scan2map_batch(
            self.h, C.c_int(batch["B"]), batch["carr"], batch["sarr"],
            poses_inout.ctypes.data_as(C.POINTER(C.c_double)), None))

    def scan2map_submit(self, batch, poses_in: np.ndarray) -> int:
        """msfl_scan2map_batch_submit on a prepared batch: enqueues upload + kernels + pose download and
        returns a ticket; up to 3 batches may be in flight (repack of k+2, upload of k+1, kernels of k)."""
        assert poses_in.dtype == np.float64 and poses_in.flags.c_contiguous
        t = C.c_int(-1)
        self._check(self.lib.msfl_scan2map_batch_submit(
            self.h, C.c_int(batch["B"]), batch["carr"], batch["sarr"],
            poses_in.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(0), C.byref(t)))
        return t.value

    def scan2map_wait(self, ticket: int, poses_out: np.ndarray):
        """msfl_scan2map_batch_wait: blocks until the batch of `ticket` is done, fills poses_out (B,7)."""
        assert poses_out.dtype == np.float64 and poses_out.flags.c_contiguous
        return self._check(self.lib.msfl_scan2map_batch_wait(
            self.h, C.c_int(ticket), poses_out.ctypes.data_as(C.POINTER(C.c_double)), None))

    def scan2map_batch_device(self, B, d_corner, d_corner_off, n_corner_total, d_surf, d_surf_off,
                              n_surf_total, d_poses, d_stats=0):
        """All arguments are raw device pointers (ints); enqueues and returns without syncing."""
        self._check(self.lib.msfl_scan2map_batch_device(
            self.h, C.c_int(B), C.c_void_p(d_corner), C.c_void_p(d_corner_off), C.c_size_t(n_corner_total),
            C.c_void_p(d_surf), C.c_void_p(d_surf_off), C.c_size_t(n_surf_total), C.c_void_p(d_poses),
            C.c_void_p(d_stats) if d_stats else None))

    def associate_map(self, scan_corner, scan_surf, pose):
        vc, vs = _View(scan_corner), _View(scan_surf)
        n = vc.n + vs.n
        knn = np.full((n, 5), -1, np.int32)
        corr = np.zeros((n, 6), np.float64)
        x = _pose(pose)
        self._check(self.lib.msfl_associate_map(self.h, C.byref(vc.cloud), C.byref(vs.cloud),
                                                x.ctypes.data_as(C.POINTER(C.c_double)),
                                                knn.ctypes.data_as(C.POINTER(C.c_int32)),
                                                corr.ctypes.data_as(C.POINTER(C.c_double))))
        return knn, corr

    def accumulate(self, p_xyz, corr6, n_edge, n_plane, pose):
        p = np.ascontiguousarray(p_xyz, dtype=np.float32).reshape(-1, 3)
        c = np.ascontiguousarray(corr6, dtype=np.float64).reshape(-1, 6)
        x = _pose(pose)
        cost = C.c_double()
        H = np.zeros((6, 6))
        g = np.zeros(6)
        self._check(self.lib.msfl_accumulate(self.h, p.ctypes.data_as(C.POINTER(C.c_float)),
                                             c.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(n_edge),
                                             C.c_int(n_plane), x.ctypes.data_as(C.POINTER(C.c_double)),
                                             C.byref(cost), H.ctypes.data_as(C.POINTER(C.c_double)),
                                             g.ctypes.data_as(C.POINTER(C.c_double))))
        return cost.value, H, g

    # ---------------------------------------------------------------- scan-to-scan
    def scan2scan(self, last_corner, last_surf, curr_sharp, curr_flat, pose, want_stats=True):
        views = [_View(a) for a in (last_corner, last_surf, curr_sharp, curr_flat)]
        x = _pose(pose)
        st = Stats() if want_stats else None
        rc = self._check(self.lib.msfl_scan2scan(self.h, *[C.byref(v.cloud) for v in views],
                                                 x.ctypes.data_as(C.POINTER(C.c_double)),
                                                 C.byref(st) if st is not None else None))
        return rc, x, (st.as_dict() if st is not None else None)

    def scan2scan_batch(self, last_corners, last_surfs, curr_sharps, curr_flats, poses, want_stats=False):
        """B independent MatchScan2Scan problems in one call; returns (status[B], poses[B,7], stats list or None)."""
        B = len(last_corners)
        views = [[_View(a) for a in lst] for lst in (last_corners, last_surfs, curr_sharps, curr_flats)]
        arrs = [(Cloud * B)(*[v.cloud for v in vs]) for vs in views]
        x = np.ascontiguousarray(poses, dtype=np.float64).reshape(B, 7).copy()
        status = np.zeros(B, np.int32)
        st = (Stats * B)() if want_stats else None
        self._check(self.lib.msfl_scan2scan_batch(self.h, C.c_int(B), *arrs, x.ctypes.data_as(C.POINTER(C.c_double)),
                                                  status.ctypes.data_as(C.POINTER(C.c_int32)), st))
        return status, x, ([s.as_dict() for s in st] if st is not None else None)

    def associate_scan(self, last_corner, last_surf, curr_sharp, curr_flat, pose):
        views = [_View(a) for a in (last_corner, last_surf, curr_sharp, curr_flat)]
        x = _pose(pose)
        assoc = np.full(2 * views[2].n + 3 * views[3].n + 1, -1, np.int32)
        self._check(self.lib.msfl_associate_scan(self.h, *[C.byref(v.cloud) for v in views],
                                                 x.ctypes.data_as(C.POINTER(C.c_double)),
                                                 assoc.ctypes.data_as(C.POINTER(C.c_int32))))
        return assoc[:-1]

    # ---------------------------------------------------------------- extraction / voxel grid
    def extract_features(self, raw, ring=None, T_lidar2imu=None):
        v = _View(raw, ring)
        n = v.n
        full = np.zeros((n, 4), np.float32)
        fring = np.zeros(n, np.uint16)
        curv = np.zeros(n, np.float32)
        label = np.zeros(n, np.int32)
        idx = [np.zeros(n, np.int32) for _ in range(4)]
        f = Features()
        f.full_xyzi = full.ctypes.data_as(C.POINTER(C.c_float))
        f.full_ring = fring.ctypes.data_as(C.POINTER(C.c_uint16))
        f.curvature = curv.ctypes.data_as(C.POINTER(C.c_float))
        f.label = label.ctypes.data_as(C.POINTER(C.c_int32))
        f.idx_sharp, f.idx_less_sharp, f.idx_flat, f.idx_less_flat = [
            a.ctypes.data_as(C.POINTER(C.c_int32)) for a in idx]
        T = _pose(T_lidar2imu) if T_lidar2imu is not None else None
        self._check(self.lib.msfl_extract_features(
            self.h, C.byref(v.cloud), T.ctypes.data_as(C.POINTER(C.c_double)) if T is not None else None,
            C.byref(f)))
        nf = f.n_full
        return {
            "full": full[:nf], "ring": fring[:nf], "curvature": curv[:nf], "label": label[:nf],
            "idx_sharp": idx[0][: f.n_sharp], "idx_less_sharp": idx[1][: f.n_less_sharp],
            "idx_flat": idx[2][: f.n_flat], "idx_less_flat": idx[3][: f.n_less_flat],
        }

    def extract_features_batch(self, raws, rings, T_lidar2imu=None):
        """msfl_extract_features_batch: B raw scans through one launch sequence; returns a list of dicts like
        extract_features."""
        B = len(raws)
        views = [_View(r, g) for r, g in zip(raws, rings)]
        carr = (Cloud * B)(*[v.cloud for v in views])
        feats = (Features * B)()
        bufs = []
        for b, v in enumerate(views):
            n = v.n
            full, fring, curv, label = np.zeros((n, 4), np.float32), np.zeros(n, np.uint16), np.zeros(n, np.float32), np.zeros(n, np.int32)
            idx = [np.zeros(n, np.int32) for _ in range(4)]
            f = feats[b]
            f.full_xyzi = full.ctypes.data_as(C.POINTER(C.c_float))
            f.full_ring = fring.ctypes.data_as(C.POINTER(C.c_uint16))
            f.curvature = curv.ctypes.data_as(C.POINTER(C.c_float))
            f.label = label.ctypes.data_as(C.POINTER(C.c_int32))
            f.idx_sharp, f.idx_less_sharp, f.idx_flat, f.idx_less_flat = [a.ctypes.data_as(C.POINTER(C.c_int32)) for a in idx]
            bufs.append((full, fring, curv, label, idx))
        T = _pose(T_lidar2imu) if T_lidar2imu is not None else None
        self._check(self.lib.msfl_extract_features_batch(
            self.h, C.c_int(B), carr, T.ctypes.data_as(C.POINTER(C.c_double)) if T is not None else None, feats))
        out = []
        for b, (full, fring, curv, label, idx) in enumerate(bufs):
            f = feats[b]
            nf = f.n_full
            out.append({"full": full[:nf], "ring": fring[:nf], "curvature": curv[:nf], "label": label[:nf],
                        "idx_sharp": idx[0][: f.n_sharp], "idx_less_sharp": idx[1][: f.n_less_sharp],
                        "idx_flat": idx[2][: f.n_flat], "idx_less_flat": idx[3][: f.n_less_flat]})
        return out

    def prepare_raw_batch(self, raws, rings):
        """msfl_cloud table of B raw scans in the reference's PointXYZIRT layout (built once, reused by timed loops)."""
        views = [_View(r, g) for r, g in zip(raws, rings)]
        return {"B": len(views), "views": views, "arr": (Cloud * len(views))(*[v.cloud for v in views])}

    def register_and_match_batch(self, raw_batch, poses, T_lidar2imu=None, leaf_corner=0.2, leaf_surf=0.4, want_counts=False):
        """msfl_register_and_match_batch: raw scans -> poses (registration, VoxelGrid, scan-to-map on the GPU)."""
        if not isinstance(raw_batch, dict):
            raw_batch = self.prepare_raw_batch(*raw_batch)
        B = raw_batch["B"]
        x = np.ascontiguousarray(poses, dtype=np.float64).reshape(B, 7).copy()
        T = _pose(T_lidar2imu) if T_lidar2imu is not None else None
        cnt = (_lib.ChainCounts * B)() if want_counts else None
        self._check(self.lib.msfl_register_and_match_batch(
            self.h, C.c_int(B), raw_batch["arr"], T.ctypes.data_as(C.POINTER(C.c_double)) if T is not None else None,
            C.c_float(leaf_corner), C.c_float(leaf_surf), x.ctypes.data_as(C.POINTER(C.c_double)), cnt, None))
        if want_counts:
            return x, [{k: getattr(c, k) for k, _ in _lib.ChainCounts._fields_} for c in cnt]
        return x

    def replay_batch(self, raw_batch, odom, poses, compose=True, T_lidar2imu=None, leaf_corner=0.2, leaf_surf=0.4):
        """msfl_replay_batch: B consecutive raw scans -> odometry of the B - 1 pairs + scan-to-map poses.
        odom (B,7): entry b >= 1 = initial guess of pose_curr2last of scan b.  Returns (odom, odom_status, poses)."""
        if not isinstance(raw_batch, dict):
            raw_batch = self.prepare_raw_batch(*raw_batch)
        B = raw_batch["B"]
        od = np.ascontiguousarray(odom, dtype=np.float64).reshape(B, 7).copy()
        x = np.ascontiguousarray(poses, dtype=np.float64).reshape(B, 7).copy()
        status = np.zeros(B, np.int32)
        T = _pose(T_lidar2imu) if T_lidar2imu is not None else None
        self._check(self.lib.msfl_replay_batch(
            self.h, C.c_int(B), raw_batch["arr"], T.ctypes.data_as(C.POINTER(C.c_double)) if T is not None else None,
            C.c_float(leaf_corner), C.c_float(leaf_surf), od.ctypes.data_as(C.POINTER(C.c_double)),
            status.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(1 if compose else 0),
            x.ctypes.data_as(C.POINTER(C.c_double)), None, None))
        return od, status, x

    def voxel_grid(self, xyzi, leaf):
        v = _View(xyzi)
        out = np.zeros((max(v.n, 1), 4), np.float32)
        n_out = C.c_size_t(0)
        self._check(self.lib.msfl_voxel_grid(self.h, C.byref(v.cloud), C.c_float(leaf),
                                             out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n_out)))
        return out[: n_out.value].copy()


class HybridGrid:
    """GPU-resident STGM map of one feature class (hybrid_grid.h:27-39)."""

    def __init__(self, engine: Engine, resolution: float = 3.0, leaf: float = 0.4):
        self.engine = engine
        h = C.c_void_p()
        engine._check(engine.lib.msfl_map_create(engine.h, C.c_float(resolution), C.c_float(leaf), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.engine.lib.msfl_map_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if self.engine.h:
                self.close()
        except Exception:
            pass

    def InsertScan(self, scan, pose=None):
        """scan already in the world frame (pose None) or sensor-frame scan + pose_map_scan2world."""
        v = _View(scan)
        x = _pose(pose) if pose is not None else None
        self.engine._check(self.engine.lib.msfl_map_insert(
            self.h, C.byref(v.cloud), x.ctypes.data_as(C.POINTER(C.c_double)) if x is not None else None))

    def size(self):
        n, c = C.c_size_t(), C.c_size_t()
        self.engine._check(self.engine.lib.msfl_map_size(self.h, C.byref(n), C.byref(c)))
        return n.value, c.value

    def _download(self, which, n):
        out = np.zeros((max(n, 1), 4), np.float32)
        got = C.c_size_t()
        self.engine._check(self.engine.lib.msfl_map_download(self.h, C.c_int(which), out.ctypes.data_as(C.POINTER(C.c_float)),
                                                             C.c_size_t(out.shape[0]), C.byref(got)))
        return out[: got.value].copy()

    def GetSurroundedCloud(self, scan, pose, download=True):
        v = _View(scan)
        x = _pose(pose)
        n = C.c_size_t()
        self.engine._check(self.engine.lib.msfl_map_surround(self.h, C.byref(v.cloud),
                                                             x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n)))
        return self._download(0, n.value) if download else n.value

    def dump(self):
        return self._download(1, self.size()[0])


def set_submap_from_maps(engine: Engine, corner: "HybridGrid", surf: "HybridGrid"):
    engine._check(engine.lib.msfl_set_submap_from_maps(engine.h, corner.h, surf.h))


def mapping_frame(engine: Engine, corner: "HybridGrid", surf: "HybridGrid", cloud_corner_less_sharp, cloud_surf_less_flat, pose,
                  want_stats=True):
    """msfl_mapping_frame: one frame of LaserMapping (laser_mapping.cc:258-340) -- VoxelGrid of the scan's feature clouds,
    GetSurroundedCloud of both maps, the size gate, MatchScan2Map, InsertScan at the refined pose -- in one call.
    Returns (matched, pose, stats dict or None)."""
    vc, vs = _View(cloud_corner_less_sharp), _View(cloud_surf_less_flat)
    x = _pose(pose)
    st = Stats() if want_stats else None
    matched = C.c_int32(0)
    engine._check(engine.lib.msfl_mapping_frame(engine.h, corner.h, surf.h, C.byref(vc.cloud), C.byref(vs.cloud),
                                                x.ctypes.data_as(C.POINTER(C.c_double)), C.byref(matched),
                                                C.byref(st) if st is not None else None))
    return bool(matched.value), x, (st.as_dict() if st is not None else None)


class TimestampedPointCloud:
    """The five clouds of the reference's TimestampedPointCloud (timestamped_pointcloud.h:11-48);
    each member is an (n,4) float32 array (plus ``*_ring`` uint16 for PointXYZIRT clouds)."""

    def __init__(self, **kw):
        self.time = 0
        self.cloud_full_res = None
        self.cloud_corner_sharp = None
        self.cloud_corner_less_sharp = None
        self.cloud_surf_flat = None
        self.cloud_surf_less_flat = None
        self.ring_corner_less_sharp = None
        self.ring_surf_less_flat = None
        for k, v in kw.items():
            setattr(self, k, v)


class MappingScanMatcher:
    """mapping_scan_matcher.h:12-22.  LiDAR-only branch (is_initialized == False)."""

    def __init__(self, engine: Engine | None = None, **engine_kw):
        self.engine = engine or Engine(**engine_kw)
        self.last_stats = None

    def MatchScan2Map(self, cloud_map: TimestampedPointCloud, scan_curr: TimestampedPointCloud,
                      is_initialized: bool, pose_estimate_map_scan2world, preintegration=None,
                      gravity_vector=None, velocity=None):
        """Returns (True, pose) like the reference (the bool is always true, :277).  With
        is_initialized the deskew factors are used; `preintegration` = (sum_dt, delta_q, delta_p) buffers
        and the pose must already include the caller's IMU-only predict (:35-60)."""
        self.engine.set_submap(cloud_map.cloud_corner_less_sharp, cloud_map.cloud_surf_less_flat)
        if is_initialized:
            sum_dt, dq, dp = preintegration
            _, pose, st = self.engine.scan2map_deskew(scan_curr.cloud_corner_less_sharp, scan_curr.cloud_surf_less_flat,
                                                      sum_dt, dq, dp, velocity, gravity_vector,
                                                      pose_estimate_map_scan2world)
            self.last_stats = st
            return True, pose
        _, pose, st = self.engine.scan2map(scan_curr.cloud_corner_less_sharp, scan_curr.cloud_surf_less_flat,
                                           pose_estimate_map_scan2world)
        self.last_stats = st
        return True, pose


class OdometryScanMatcher:
    """odometry_scan_matcher.h:8-13."""

    def __init__(self, engine: Engine | None = None, **engine_kw):
        self.engine = engine or Engine(**engine_kw)
        self.last_stats = None

    def MatchScan2Scan(self, scan_last: TimestampedPointCloud, scan_curr: TimestampedPointCloud,
                       pose_estimate_curr2last):
        """Returns (ok, pose); ok is False when fewer than 10 correspondences were found
        (odometry_scan_matcher.cc:262-267)."""
        rc, pose, st = self.engine.scan2scan(
            to_pcl(scan_last.cloud_corner_less_sharp, scan_last.ring_corner_less_sharp),
            to_pcl(scan_last.cloud_surf_less_flat, scan_last.ring_surf_less_flat),
            scan_curr.cloud_corner_sharp, scan_curr.cloud_surf_flat, pose_estimate_curr2last)
        self.last_stats = st
        return rc == MSFL_OK, pose


class ScanRegistration:
    """The feature-extraction block of RealHandleLaserCloudMessage (msf_loam_node.cc:160-371)."""

    def __init__(self, engine: Engine | None = None, lidar2imu=None, **engine_kw):
        self.engine = engine or Engine(**engine_kw)
        self.lidar2imu = lidar2imu

    def extract(self, xyzi, ring) -> TimestampedPointCloud:
        f = self.engine.extract_features(xyzi, ring, self.lidar2imu)
        full, rg = f["full"], f["ring"]
        return TimestampedPointCloud(
            cloud_full_res=full,
            cloud_corner_sharp=full[f["idx_sharp"]],
            cloud_corner_less_sharp=full[f["idx_less_sharp"]],
            cloud_surf_flat=full[f["idx_flat"]],
            cloud_surf_less_flat=full[f["idx_less_flat"]],
            ring_corner_less_sharp=rg[f["idx_less_sharp"]],
            ring_surf_less_flat=rg[f["idx_less_flat"]],
            features=f,
        )
