"""dev: latency of the STGM map calls and VoxelGrid through the host API."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msf_loam_b200 import Engine, HybridGrid, default_params
from msf_loam_b200 import synth as S
def t(fn, n=20):
    for _ in range(3): fn()
    t0 = time.perf_counter()
    for _ in range(n): r = fn()
    return (time.perf_counter() - t0) / n * 1e6, r
scene = S.make_scene(); traj = S.trajectory(12)
e = Engine(default_params())
feats = [e.extract_features(*S.raycast_scan(scene, "vlp16", traj[k], seed=100 + k)) for k in range(12)]
gc, gs = HybridGrid(e, 3.0, 0.2), HybridGrid(e, 3.0, 0.4)
for k in range(10):
    f = feats[k]
    gc.InsertScan(S.transform_cloud(traj[k], f["full"][f["idx_less_sharp"]]))
    gs.InsertScan(S.transform_cloud(traj[k], f["full"][f["idx_less_flat"]]))
f = feats[10]
c = S.transform_cloud(traj[10], f["full"][f["idx_less_sharp"]]); s = S.transform_cloud(traj[10], f["full"][f["idx_less_flat"]])
us_ic, _ = t(lambda: gc.InsertScan(c)); us_is, _ = t(lambda: gs.InsertScan(s))
us_sc, rc = t(lambda: gc.GetSurroundedCloud(f["full"][f["idx_less_sharp"]], traj[10], download=False))
us_ss, rs = t(lambda: gs.GetSurroundedCloud(f["full"][f["idx_less_flat"]], traj[10], download=False))
us_v, _ = t(lambda: e.voxel_grid(s, 0.4))
print(f"map sizes {gc.size()} / {gs.size()} | InsertScan corner({len(c)}) {us_ic:.0f} us, surf({len(s)}) {us_is:.0f} us | "
      f"GetSurroundedCloud (device) corner {us_sc:.0f} us, surf {us_ss:.0f} us | voxel_grid({len(s)}) {us_v:.0f} us")
