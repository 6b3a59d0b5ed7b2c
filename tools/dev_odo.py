"""dev: scan-to-scan + extraction latency at the C-ABI level (inputs pre-packed)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msf_loam_b200 import Engine, default_params
from msf_loam_b200 import synth as S
from msf_loam_b200.engine import to_pcl
sensor = sys.argv[1] if len(sys.argv) > 1 else "vlp16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
scene = S.make_scene("room40" if sensor == "vlp16" else "room80"); traj = S.trajectory(3)
scans = [S.raycast_scan(scene, sensor, traj[k], seed=100 + k) for k in range(2)]
e = Engine(default_params())
f = [e.extract_features(x, r, None) for x, r in scans]
def cl(ff, key): return to_pcl(ff["full"][ff[key]], ff["ring"][ff[key]])
lc, ls, cs, cf = cl(f[0], "idx_less_sharp"), cl(f[0], "idx_less_flat"), cl(f[1], "idx_sharp"), cl(f[1], "idx_flat")
print(sensor, "last less_sharp", len(lc), "less_flat", len(ls), "sharp", len(cs), "flat", len(cf))
for _ in range(3): e.scan2scan(lc, ls, cs, cf, S.pose_identity(), want_stats=False)
t0 = time.perf_counter()
for _ in range(n): rc, x, _ = e.scan2scan(lc, ls, cs, cf, S.pose_identity(), want_stats=False)
print("scan2scan us/call", (time.perf_counter() - t0) / n * 1e6, "rc", rc)
x0, r0 = scans[1]
t0 = time.perf_counter()
for _ in range(n): e.extract_features(x0, r0, None)
print("extract us/call", (time.perf_counter() - t0) / n * 1e6)
# raw C-ABI call with preallocated outputs (no Python marshalling inside the timed loop)
import ctypes as C
from msf_loam_b200._lib import Features
from msf_loam_b200.engine import _View
v = _View(to_pcl(x0, r0))
nn_ = v.n
full = np.zeros((nn_, 4), np.float32); fring = np.zeros(nn_, np.uint16); idx = [np.zeros(nn_, np.int32) for _ in range(4)]
ft = Features()
ft.full_xyzi = full.ctypes.data_as(C.POINTER(C.c_float)); ft.full_ring = fring.ctypes.data_as(C.POINTER(C.c_uint16))
ft.idx_sharp, ft.idx_less_sharp, ft.idx_flat, ft.idx_less_flat = [a.ctypes.data_as(C.POINTER(C.c_int32)) for a in idx]
for _ in range(3): e.lib.msfl_extract_features(e.h, C.byref(v.cloud), None, C.byref(ft))
t0 = time.perf_counter()
for _ in range(n): e.lib.msfl_extract_features(e.h, C.byref(v.cloud), None, C.byref(ft))
print("extract (C ABI, preallocated outputs: full cloud + rings + 4 index lists) us/call", (time.perf_counter() - t0) / n * 1e6)
e.set_profiling(True)
