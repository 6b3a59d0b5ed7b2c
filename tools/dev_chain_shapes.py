"""Dev: chain (raw -> pose) and replay timing with the adaptive k_feat_pick shape vs the big shape forced."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from conftest import make_map_case
from msf_loam_b200 import Engine, default_params, synth as S

case = make_map_case()
qs = case["queries"]
B = 256
tri = [0, 1, 2, 1] * (B // 4)
for forced in (None, "2"):
    if forced is None: os.environ.pop("MSFL_PICK_SHAPE", None)
    else: os.environ["MSFL_PICK_SHAPE"] = forced
    e = Engine(default_params())
    e.set_submap(case["map_corner"], case["map_surf"])
    batch = e.prepare_raw_batch([qs[t]["raw"][0].copy() for t in tri], [qs[t]["raw"][1].copy() for t in tri])
    inits = np.stack([qs[t]["init"] for t in tri])
    ident = np.tile(S.pose_identity(), (B, 1))
    for _ in range(2):
        e.register_and_match_batch(batch, inits); e.replay_batch(batch, ident, inits, compose=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): e.register_and_match_batch(batch, inits)
    t1 = time.perf_counter()
    for _ in range(5): e.replay_batch(batch, ident, inits, compose=False)
    t2 = time.perf_counter()
    print("pick shape", forced or "adaptive", "chain %.2f ms  replay %.2f ms per 256 scans" % ((t1 - t0) / 5 * 1e3, (t2 - t1) / 5 * 1e3))
    e.close()
