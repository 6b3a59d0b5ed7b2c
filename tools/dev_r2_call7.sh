mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c7_pytest.log
cat gpurun_out/r2c7_pytest.log
python - <<'PY'
# where does the single-scan call spend its time: kernel (CUDA events) vs host + copies
import time, numpy as np, sys
sys.path.insert(0,'.')
import bench
from msf_loam_b200 import Engine, default_params, synth as S
traj, scans = bench.raw_scans('vlp16', 1, 8)
e0=Engine(default_params(**bench.OVER))
mc, ms, queries, _ = bench.build_case(lambda x, r: e0.extract_features(x, r, None), e0.voxel_grid, 'vlp16', traj, scans)
e0.close()
c0, s0, gt = queries[0]
init = S.perturb_pose(gt, np.random.default_rng(1))
for G in (0, 8, 16):
    e1 = Engine(default_params(lm_cluster=G, **bench.OVER)); e1.set_submap(mc, ms)
    for _ in range(5): e1.scan2map(c0, s0, init, want_stats=False)
    e1.get_profile(); e1.set_profiling(True)
    t0=time.perf_counter()
    for _ in range(50): e1.scan2map(c0, s0, init, want_stats=False)
    t=(time.perf_counter()-t0)/50*1e6
    ms_, cnt = e1.get_profile()
    print('G', G, 'wall us/scan %.1f'%t, 'stage ms per call', [round(m/50*1e3,1) for m in ms_], cnt)
    e1.close()
PY
