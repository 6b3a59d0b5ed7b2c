"""Wire / on-disk formats of the clouds that reach the hot path.

* ``pointcloud2_view``  -- a ``sensor_msgs/PointCloud2`` data buffer viewed as an ``msfl_cloud`` (the
  field mapping ``pcl::fromROSMsg`` performs at msf_loam_node.cc:166-167 for the fields registered in
  common.h:53-62).  No copy: the float4 unpack runs on the GPU inside ``msfl_extract_features``.
* ``read_kitti_bin``    -- KITTI odometry ``velodyne/NNNNNN.bin`` (float32 x, y, z, reflectance), the
  reader of kitti_helper.cc:21-32 / :145-154.  KITTI files carry no ring; like the reference
  (":todo write scan ring here", kitti_helper.cc:152) ring assignment is left to the caller --
  ``rings_from_elevation`` offers the usual nearest-beam assignment for replay.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import _View

PC2_FLOAT32, PC2_UINT16 = 7, 4


def pointcloud2_view(data, width, height, point_step, row_step, fields, is_bigendian=False):
    """fields: iterable of (name, offset, datatype, count).  Returns an object usable wherever the
    engine accepts a cloud (it keeps `data` alive)."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
    lib = _lib.load_library()
    arr = (_lib.Pc2Field * len(fields))(*[_lib.Pc2Field(n.encode(), o, t, c) for (n, o, t, c) in fields])
    cloud = _lib.Cloud()
    rc = lib.msfl_cloud_from_pointcloud2(buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint32(width), C.c_uint32(height),
                                         C.c_uint32(point_step), C.c_uint32(row_step), C.c_int(1 if is_bigendian else 0),
                                         arr, C.c_int(len(fields)), C.byref(cloud))
    if rc < 0:
        raise _lib.MsflError(rc, lib.msfl_last_error().decode())
    v = _View.__new__(_View)
    v.buf, v.cloud, v.aux = buf, cloud, arr
    return v


def make_pointcloud2(xyzi, ring, point_step=22, offsets=(0, 4, 8, 12, 16)):
    """Test/bench helper: serialise arrays into a PointCloud2-style buffer (default: the packed 22-byte
    layout with unaligned fields some drivers emit).  Returns (bytes, fields)."""
    xyzi = np.asarray(xyzi, dtype=np.float32).reshape(-1, 4)
    ring = np.asarray(ring, dtype=np.uint16)
    n = xyzi.shape[0]
    raw = np.zeros((n, point_step), np.uint8)
    ox, oy, oz, oi, orr = offsets
    for col, off in zip(range(4), (ox, oy, oz, oi)):
        raw[:, off:off + 4] = xyzi[:, col].copy().view(np.uint8).reshape(n, 4)
    raw[:, orr:orr + 2] = ring.copy().view(np.uint8).reshape(n, 2)
    fields = [("x", ox, PC2_FLOAT32, 1), ("y", oy, PC2_FLOAT32, 1), ("z", oz, PC2_FLOAT32, 1),
              ("intensity", oi, PC2_FLOAT32, 1), ("ring", orr, PC2_UINT16, 1)]
    return raw.tobytes(), fields


def read_kitti_bin(path):
    """(n, 4) float32: x, y, z, reflectance (kitti_helper.cc:21-32, :145-150)."""
    a = np.fromfile(path, dtype=np.float32)
    return a[: (a.size // 4) * 4].reshape(-1, 4)


def rings_from_elevation(xyz, n_rings=64, fov_up_deg=2.0, fov_down_deg=-24.8):
    """Nearest-beam ring index from the vertical angle (ring increases with elevation, README.md:56-58)."""
    xyz = np.asarray(xyz, dtype=np.float64)
    el = np.degrees(np.arctan2(xyz[:, 2], np.hypot(xyz[:, 0], xyz[:, 1])))
    r = np.rint((el - fov_down_deg) / (fov_up_deg - fov_down_deg) * (n_rings - 1))
    return np.clip(r, 0, n_rings - 1).astype(np.uint16)
