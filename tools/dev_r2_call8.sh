mkdir -p gpurun_out
cat > /tmp/lat.py <<'PY'
import time, numpy as np, sys, os, struct, subprocess
sys.path.insert(0,'.')
import bench
from msf_loam_b200 import Engine, default_params, synth as S
import oracle as O
traj, scans = bench.raw_scans('vlp16', 2, 8)
e0=Engine(default_params(**bench.OVER))
mc, ms, queries, _ = bench.build_case(lambda x, r: e0.extract_features(x, r, None), e0.voxel_grid, 'vlp16', traj, scans)
e0.close()
c0, s0, gt = queries[0]
init = S.perturb_pose(gt, np.random.default_rng(1))
for G in (8, 16):
    e1 = Engine(default_params(lm_cluster=G)); e1.set_submap(mc, ms)
    for _ in range(5): e1.scan2map(c0, s0, init, want_stats=False)
    e1.get_profile(); e1.set_profiling(True)
    t0=time.perf_counter()
    for _ in range(50): e1.scan2map(c0, s0, init, want_stats=False)
    t=(time.perf_counter()-t0)/50*1e6
    ms_, cnt = e1.get_profile()
    print(os.environ.get('MSFL_LIB_PATH','tree'), 'G', G, 'wall us/scan %.1f'%t, 'kernel us', round(ms_[1]/50*1e3,1))
    e1.close()
PY
python /tmp/lat.py
MSFL_LIB_PATH=$PWD/msf_loam_b200/libmsfl_f512.so python /tmp/lat.py
# the C driver: drop-in latency without Python
python - <<'PY'
import sys, os, struct, subprocess, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import make_map_case
from msf_loam_b200 import synth as S
case = make_map_case(); q = case["queries"][0]
f0, f1 = case["queries"][0]["features"], case["queries"][1]["features"]
lc, rlc = f0["full"][f0["idx_less_sharp"]], f0["ring"][f0["idx_less_sharp"]]
ls, rls = f0["full"][f0["idx_less_flat"]], f0["ring"][f0["idx_less_flat"]]
cs, cf = f1["full"][f1["idx_sharp"]], f1["full"][f1["idx_flat"]]
arrays = [case["map_corner"], case["map_surf"], q["corner"], q["surf"], lc, ls, cs, cf]
with open('/tmp/case.bin','wb') as f:
    f.write(struct.pack("8i", *[a.shape[0] for a in arrays]))
    for a in arrays: f.write(np.ascontiguousarray(a, dtype=np.float32).tobytes())
    f.write(rlc.astype(np.float32).tobytes()); f.write(rls.astype(np.float32).tobytes())
    f.write(np.asarray(q["init"], dtype=np.float64).tobytes()); f.write(S.pose_identity().astype(np.float64).tobytes())
subprocess.run(["g++","-std=c++14","-O2","-Iinclude","-Imsf_loam_b200/adapter","-Itests/adapter_stubs","tests/c_abi/adapter_driver.cc","-Lmsf_loam_b200","-lmsfl","-Wl,-rpath,"+os.path.abspath("msf_loam_b200"),"-o","/tmp/adrv"],check=True)
print(subprocess.run(["/tmp/adrv","/tmp/case.bin","bench"],capture_output=True,text=True).stdout)
PY
