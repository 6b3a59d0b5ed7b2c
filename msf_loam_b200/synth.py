"""Seeded synthetic LiDAR scenes: numpy ray-casting of spinning multi-beam
sensors (VLP-16 / HDL-64E / OS1-128 beam tables) against a box room + ground + poles + boxes.

Host-side input generation only -- no part of the matching path.  Clouds come out in the
reference's input contract (README.md:56-58, scripts/validate_velodyne_cloud_in_bag.py:1-15):
``ring`` increases with elevation, points are clockwise within a ring; the raw cloud is in
firing order (azimuth-major) so that the ring split (msf_loam_node.cc:128-156) has work to do.
Poses are ``[tx ty tz qx qy qz qw]`` (rigid_transform.h:59-64).
"""
from __future__ import annotations

import numpy as np

SENSORS = {
    # name: (elevations in degrees ascending, azimuth steps)
    "vlp16": (np.arange(-15.0, 15.1, 2.0), 1812),
    "hdl64": (np.linspace(-24.8, 2.0, 64), 2030),
    "os1-128": (np.linspace(-22.5, 22.5, 128), 2048),
}


# ----------------------------------------------------------------------------- pose algebra
def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


def quat_to_R(q):
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def rotvec_to_quat(v):
    v = np.asarray(v, dtype=np.float64)
    th = np.linalg.norm(v)
    if th < 1e-12:
        return np.array([0.5 * v[0], 0.5 * v[1], 0.5 * v[2], 1.0])
    s = np.sin(0.5 * th) / th
    return np.array([s * v[0], s * v[1], s * v[2], np.cos(0.5 * th)])


def pose_identity():
    return np.array([0, 0, 0, 0, 0, 0, 1.0])


def pose_mul(a, b):
    """Rigid3 operator* (rigid_transform.h:105-111): t = Ra tb + ta, q = (qa qb).normalized()."""
    q = quat_mul(a[3:], b[3:])
    q /= np.linalg.norm(q)
    t = quat_to_R(a[3:]) @ b[:3] + a[:3]
    return np.concatenate([t, q])


def pose_inv(a):
    qc = np.array([-a[3], -a[4], -a[5], a[6]])
    return np.concatenate([-(quat_to_R(qc) @ a[:3]), qc])


def pose_error(a, b):
    """(translation error [m], rotation angle of qa^-1 qb [rad])."""
    dt = float(np.linalg.norm(np.asarray(a[:3]) - np.asarray(b[:3])))
    qa = np.asarray(a[3:], dtype=np.float64)
    qb = np.asarray(b[3:], dtype=np.float64)
    qd = quat_mul(np.array([-qa[0], -qa[1], -qa[2], qa[3]]), qb)
    ang = 2.0 * np.arctan2(np.linalg.norm(qd[:3]), abs(qd[3]))
    return dt, float(ang)


def perturb_pose(pose, rng, trans=0.10, rot_deg=1.0):
    """pose o (random translation of norm `trans`, random-axis rotation of `rot_deg`)."""
    d = rng.normal(size=3)
    d *= trans / np.linalg.norm(d)
    ax = rng.normal(size=3)
    ax *= np.deg2rad(rot_deg) / np.linalg.norm(ax)
    return pose_mul(pose, np.concatenate([d, rotvec_to_quat(ax)]))


def transform_cloud(pose, xyzi):
    """TransformPointCloud semantics (rigid_transform.h:140-145): fp32 -> fp64 -> fp32."""
    out = np.array(xyzi, dtype=np.float32, copy=True)
    R = quat_to_R(np.asarray(pose[3:], dtype=np.float64))
    out[:, :3] = (out[:, :3].astype(np.float64) @ R.T + np.asarray(pose[:3])).astype(np.float32)
    return out


# ----------------------------------------------------------------------------- scene
def make_scene(kind="room40", seed=7):
    """40x30x6 m room (80x60x8 for the big sensors) with 12 poles and 8 boxes."""
    rng = np.random.default_rng(seed)
    if kind == "room40":
        half = np.array([20.0, 15.0])
        height = 6.0
    elif kind == "room80":
        half = np.array([40.0, 30.0])
        height = 8.0
    elif kind == "hall300":  # BASELINE config 5: a 50-scan drive whose submap reaches ~1 M points
        half = np.array([150.0, 60.0])
        height = 12.0
    else:
        raise ValueError(kind)
    n_poles, n_boxes = (12, 8) if kind != "hall300" else (60, 40)
    box_max = 3.0 if kind != "hall300" else 8.0
    poles = []
    while len(poles) < n_poles:
        c = rng.uniform(-half + 2.0, half - 2.0)
        if np.linalg.norm(c - np.array([-8.0, -3.0])) < 3.0 or (kind == "hall300" and abs(c[1] + 3.0) < 2.0):
            continue
        poles.append((c[0], c[1], 0.1))
    boxes = []
    while len(boxes) < n_boxes:
        size = rng.uniform(1.0, box_max, size=3)
        c = rng.uniform(-half + 3.0, half - 3.0)
        if abs(c[1] + 3.0) < 3.5 + (box_max - 3.0) / 2 and c[0] > -half[0] + 8:  # keep the sensor corridor clear
            continue
        boxes.append((c[0] - size[0] / 2, c[1] - size[1] / 2, 0.0,
                      c[0] + size[0] / 2, c[1] + size[1] / 2, size[2]))
    return {"half": half, "height": height, "poles": np.array(poles), "boxes": np.array(boxes)}


def trajectory(n, seed=11, step=0.5, yaw_deg=1.0, jitter_t=0.02, jitter_deg=0.2,
               start=(-8.0, -3.0, 1.5)):
    """n poses: `step` m forward + `yaw_deg` yaw per scan with seeded jitter."""
    rng = np.random.default_rng(seed)
    poses = []
    cur = np.concatenate([np.array(start, dtype=np.float64), [0, 0, 0, 1.0]])
    for _ in range(n):
        poses.append(cur.copy())
        dt = np.array([step, 0, 0]) + rng.normal(scale=jitter_t, size=3)
        rv = np.array([0, 0, np.deg2rad(yaw_deg)]) + rng.normal(scale=np.deg2rad(jitter_deg), size=3)
        cur = pose_mul(cur, np.concatenate([dt, rotvec_to_quat(rv)]))
    return poses


def raycast_scan(scene, sensor, pose, seed=0, sigma=0.01, min_range=0.3, max_range=100.0):
    """One revolution. Returns (xyzi float32 [n,4], ring uint16 [n]) in the sensor frame."""
    elev_deg, n_az = SENSORS[sensor]
    rng = np.random.default_rng(seed)
    n_ring = len(elev_deg)
    az = -2.0 * np.pi * np.arange(n_az) / n_az  # clockwise
    el = np.deg2rad(elev_deg)
    # firing order: azimuth-major, rings inner
    AZ = np.repeat(az, n_ring)
    EL = np.tile(el, n_az)
    RING = np.tile(np.arange(n_ring, dtype=np.uint16), n_az)
    ds = np.stack([np.cos(EL) * np.cos(AZ), np.cos(EL) * np.sin(AZ), np.sin(EL)], axis=1)
    R = quat_to_R(np.asarray(pose[3:], dtype=np.float64))
    o = np.asarray(pose[:3], dtype=np.float64)
    d = ds @ R.T
    with np.errstate(divide="ignore", invalid="ignore"):
        # room (origin inside): nearest positive wall hit
        lo = np.array([-scene["half"][0], -scene["half"][1], 0.0])
        hi = np.array([scene["half"][0], scene["half"][1], scene["height"]])
        t1 = (lo - o) / d
        t2 = (hi - o) / d
        t_wall = np.where(d > 0, t2, t1)
        t_wall = np.where(np.isfinite(t_wall) & (t_wall > 0), t_wall, np.inf)
        t = t_wall.min(axis=1)
        # boxes: slab test
        for bx in scene["boxes"]:
            blo, bhi = bx[:3], bx[3:]
            ta = (blo - o) / d
            tb = (bhi - o) / d
            tn = np.minimum(ta, tb).max(axis=1)
            tf = np.maximum(ta, tb).min(axis=1)
            hit = (tn <= tf) & (tn > 0)
            t = np.where(hit & (tn < t), tn, t)
        # poles: vertical cylinders
        for (cx, cy, r) in scene["poles"]:
            ox, oy = o[0] - cx, o[1] - cy
            a = d[:, 0] ** 2 + d[:, 1] ** 2
            b = 2 * (ox * d[:, 0] + oy * d[:, 1])
            c = ox * ox + oy * oy - r * r
            disc = b * b - 4 * a * c
            tc = (-b - np.sqrt(np.where(disc >= 0, disc, np.nan))) / (2 * a)
            z = o[2] + tc * d[:, 2]
            hit = np.isfinite(tc) & (tc > 0) & (z >= 0) & (z <= scene["height"])
            t = np.where(hit & (tc < t), tc, t)
    t = t + rng.normal(scale=sigma, size=t.shape)
    keep = np.isfinite(t) & (t >= min_range) & (t <= max_range)
    pts = (ds * t[:, None])[keep]
    inten = rng.uniform(0, 255, size=pts.shape[0])
    xyzi = np.concatenate([pts, inten[:, None]], axis=1).astype(np.float32)
    return xyzi, RING[keep].copy()


# ----------------------------------------------------------------------------- host VoxelGrid
def voxel_grid_np(xyzi, leaf):
    """Host restatement of pcl::VoxelGrid<PointXYZI> as the *caller* applies it
    (laser_mapping.cc:264-270): fp32 centroid of xyz+intensity per voxel, output by ascending
    voxel index, accumulation in ascending point index."""
    xyzi = np.ascontiguousarray(xyzi, dtype=np.float32).reshape(-1, 4)
    n = xyzi.shape[0]
    if n == 0:
        return xyzi.copy()
    inv = np.float32(1.0) / np.float32(leaf)
    mn = xyzi[:, :3].min(axis=0)
    mx = xyzi[:, :3].max(axis=0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(xyzi[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    starts = np.flatnonzero(np.concatenate([[True], sidx[1:] != sidx[:-1]]))
    ends = np.concatenate([starts[1:], [n]])
    out = np.zeros((len(starts), 4), np.float32)
    pts = xyzi[order]
    # sequential fp32 accumulation (matches CentroidPoint's float accumulators)
    maxc = int((ends - starts).max())
    acc = np.zeros((len(starts), 4), np.float32)
    for k in range(maxc):
        sel = (starts + k) < ends
        acc[sel] += pts[(starts + k)[sel]]
    out[:] = acc / (ends - starts).astype(np.float32)[:, None]
    return out
