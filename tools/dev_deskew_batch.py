"""Developer timing (NOT bench.py) of msfl_scan2map_deskew_batch: wall time per call vs the GPU stage times."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from msf_loam_b200 import Engine, default_params
from msf_loam_b200 import synth as S

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
traj, scans = bench.raw_scans("vlp16", 32, 8)
eng = Engine(default_params(**bench.OVER))
mc, ms, queries, _ = bench.build_case(lambda x, r: eng.extract_features(x, r, None), eng.voxel_grid, "vlp16", traj, scans)
eng.set_submap(mc, ms)
D = len(queries)
rng = np.random.default_rng(3000)
t = np.arange(0.0, 0.105 + 1e-9, 1.0 / 400.0)
tabs = []
for b in range(B):
    omega, acc, v0 = rng.normal(scale=0.05, size=3), rng.normal(scale=0.3, size=3), rng.normal(scale=0.05, size=3)
    tabs.append((t, np.stack([S.rotvec_to_quat(omega * ti) for ti in t]), np.stack([v0 * ti + 0.5 * acc * ti * ti for ti in t]),
                 tuple(rng.normal(scale=0.2, size=3)), (0.0, 0.0, 9.81)))
corners, surfs = [queries[b % D][0] for b in range(B)], [queries[b % D][1] for b in range(B)]
inits = np.stack([S.perturb_pose(queries[b % D][2], rng) for b in range(B)])
eng.scan2map_deskew_batch(corners, surfs, tabs, inits, want_stats=False)
eng.set_profiling(True)
eng.get_profile()
t0 = time.perf_counter()
for _ in range(5):
    eng.scan2map_deskew_batch(corners, surfs, tabs, inits, want_stats=False)
dt = (time.perf_counter() - t0) / 5
ms_st, cnt = eng.get_profile()
print(f"B={B}: {dt * 1e3:.2f} ms per call; GPU stage ms per call: {[round(m / 5, 3) for m in ms_st]} counts {cnt}")
t0 = time.perf_counter()
for _ in range(5):
    eng.scan2map_batch(corners, surfs, inits) if hasattr(eng, "scan2map_batch") else None
print(f"plain branch, same clouds: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms per call")
