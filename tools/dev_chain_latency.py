"""dev: per-call latency of the whole per-scan chain through the host C ABI (single scan, real-time use):
extraction -> scan-to-scan -> VoxelGrid x2 -> scan-to-map."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msf_loam_b200 import Engine, default_params
from msf_loam_b200 import synth as S
import oracle as O

def t(fn, n=30):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    return (time.perf_counter() - t0) / n * 1e6, r

for sensor, scene_kind in (("vlp16", "room40"), ("hdl64", "room80")):
    scene = S.make_scene(scene_kind); traj = S.trajectory(7)
    scans = [S.raycast_scan(scene, sensor, traj[k], seed=100 + k) for k in range(7)]
    e = Engine(default_params())
    feats = [e.extract_features(x, r, None) for x, r in scans]
    mc = np.concatenate([S.transform_cloud(traj[k], feats[k]["full"][feats[k]["idx_less_sharp"]]) for k in range(5)])
    ms = np.concatenate([S.transform_cloud(traj[k], feats[k]["full"][feats[k]["idx_less_flat"]]) for k in range(5)])
    map_c, map_s = e.voxel_grid(mc, 0.2), e.voxel_grid(ms, 0.4)
    e.set_submap(map_c, map_s)
    x, r = scans[6]
    us_ext, f = t(lambda: e.extract_features(x, r, None))
    fl, fc = feats[5], feats[6]
    def clouds(ff, key): return ff["full"][ff[key]], ff["ring"][ff[key]]
    lc, lcr = clouds(fl, "idx_less_sharp"); ls, lsr = clouds(fl, "idx_less_flat")
    cs, csr = clouds(fc, "idx_sharp"); cf, cfr = clouds(fc, "idx_flat")
    from msf_loam_b200.engine import to_pcl
    us_odo, _ = t(lambda: e.scan2scan(to_pcl(lc, lcr), to_pcl(ls, lsr), to_pcl(cs, csr), to_pcl(cf, cfr), S.pose_identity(), want_stats=False))
    corner, surf = fc["full"][fc["idx_less_sharp"]], fc["full"][fc["idx_less_flat"]]
    us_vg, _ = t(lambda: (e.voxel_grid(corner, 0.2), e.voxel_grid(surf, 0.4)))
    sc, ss = e.voxel_grid(corner, 0.2), e.voxel_grid(surf, 0.4)
    init = S.perturb_pose(traj[6], np.random.default_rng(0))
    us_map, _ = t(lambda: e.scan2map(sc, ss, init, want_stats=False))
    us_set, _ = t(lambda: e.set_submap(map_c, map_s))
    P = O.default_params()
    t0 = time.perf_counter(); O.extract_features(P, x, r, None); c_ext = (time.perf_counter() - t0) * 1e6
    t0 = time.perf_counter(); O.scan2map(P, map_c, map_s, sc, ss, init); c_map = (time.perf_counter() - t0) * 1e6
    print(f"{sensor}: {x.shape[0]} pts | extract {us_ext:.0f} us (oracle {c_ext:.0f}) | scan2scan {us_odo:.0f} us | voxelgrid x2 {us_vg:.0f} us | "
          f"set_submap({map_c.shape[0]}+{map_s.shape[0]}) {us_set:.0f} us | scan2map({sc.shape[0]}+{ss.shape[0]}) {us_map:.0f} us (oracle {c_map:.0f})")
    e.close()
