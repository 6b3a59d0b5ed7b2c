"""Developer helper: single-scan (B = 1) scan-to-map latency through the host C ABI."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import make_map_case
from msf_loam_b200 import Engine, default_params
for sensor, scene in (("vlp16", "room40"), ("hdl64", "room80")):
    case = make_map_case(sensor, scene, 5, 100 if sensor == "vlp16" else 200)
    q = case["queries"][0]
    for G in (1, 2, 4, 8):
        e = Engine(default_params(lm_cluster=G))
        e.set_submap(case["map_corner"], case["map_surf"])
        for _ in range(5):
            e.scan2map(q["corner"], q["surf"], q["init"], want_stats=False)
        t0 = time.perf_counter()
        n = 50
        for _ in range(n):
            e.scan2map(q["corner"], q["surf"], q["init"], want_stats=False)
        dt = (time.perf_counter() - t0) / n
        e.set_profiling(True)
        for _ in range(10):
            e.scan2map(q["corner"], q["surf"], q["init"], want_stats=False)
        ms, cnt = e.get_profile()
        print(f"{sensor} queries={len(q['corner'])+len(q['surf'])} lm_cluster={G}: {dt*1e6:.0f} us per scan2map; "
              f"assoc {1e3*ms[0]/max(cnt[0],1):.0f} us/launch, LM {1e3*ms[1]/max(cnt[1],1):.0f} us/launch")
        e.close()
