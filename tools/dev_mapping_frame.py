"""Developer timing (NOT bench.py) of msfl_mapping_frame and of its steps issued as separate calls (warm buffers)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle as O
from msf_loam_b200 import Engine, HybridGrid, mapping_frame, set_submap_from_maps
from msf_loam_b200 import synth as S

N = 24
P = O.default_params()
scene, traj = S.make_scene(), S.trajectory(N)
e = Engine()
scans = []
for k in range(N):
    f = e.extract_features(*S.raycast_scan(scene, "vlp16", traj[k], seed=500 + k), None)
    scans.append((f["full"][f["idx_less_sharp"]], f["full"][f["idx_less_flat"]], traj[k]))
rng = np.random.default_rng(0)
fc, fs = HybridGrid(e, 3.0, 0.2), HybridGrid(e, 3.0, 0.4)
times = []
for k, (corner, surf, gt) in enumerate(scans):
    guess = gt if k == 0 else S.perturb_pose(gt, rng)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    matched, pose, _ = mapping_frame(e, fc, fs, corner, surf, guess, want_stats=False)
    times.append(time.perf_counter() - t0)
print("frame call ms:", [round(t * 1e3, 2) for t in times], "| map sizes", fc.size(), fs.size())
print(f"frame call, frames 8..{N - 1}: {np.mean(times[8:]) * 1e3:.3f} ms per frame ({corner.shape[0]} corner / {surf.shape[0]} surf points)")
# the steps as separate calls on a second pair of maps, timed one by one
sc, ss = HybridGrid(e, 3.0, 0.2), HybridGrid(e, 3.0, 0.4)
acc = {k: 0.0 for k in ("surround", "voxel", "set_submap", "scan2map", "insert")}
for k, (corner, surf, gt) in enumerate(scans):
    guess = gt if k == 0 else S.perturb_pose(gt, rng)
    def tick(name, fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        if k >= 8: acc[name] += time.perf_counter() - t0
        return r
    n_c, n_s = tick("surround", lambda: (sc.GetSurroundedCloud(corner, guess, download=False), ss.GetSurroundedCloud(surf, guess, download=False)))
    pose = guess
    if n_c > 10 and n_s > 50:
        qc, qs = tick("voxel", lambda: (e.voxel_grid(corner, 0.2), e.voxel_grid(surf, 0.4)))
        tick("set_submap", lambda: set_submap_from_maps(e, sc, ss))
        pose = tick("scan2map", lambda: e.scan2map(qc, qs, guess, want_stats=False))[1]
    tick("insert", lambda: (sc.InsertScan(corner, pose), ss.InsertScan(surf, pose)))
print("separate calls, ms per frame:", {k: round(v / (N - 8) * 1e3, 3) for k, v in acc.items()}, "sum", round(sum(acc.values()) / (N - 8) * 1e3, 3))
