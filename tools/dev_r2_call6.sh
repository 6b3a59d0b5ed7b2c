mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c6_pytest.log
cat gpurun_out/r2c6_pytest.log
python bench.py --no-cpu --no-workloads --steps 10 > gpurun_out/r2c6_default.json 2> gpurun_out/r2c6_default.err
tail -3 gpurun_out/r2c6_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c6_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['single_scan_latency'])
PY
python - <<'PY'
# single-scan latency by cluster size, with and without per-frame submap build
import time, numpy as np, sys
sys.path.insert(0,'.')
import bench
from msf_loam_b200 import Engine, default_params, synth as S
class C: pass
cx=C(); cx.local_rank=0
cx.raw={'vlp16': bench.raw_scans('vlp16', 1, 8), 'hdl64': bench.raw_scans('hdl64', 1, 8)}
for wl in ('vlp16','hdl64'):
    traj, scans = cx.raw[wl]
    e0=Engine(default_params(**bench.OVER))
    mc, ms, queries, _ = bench.build_case(lambda x, r: e0.extract_features(x, r, None), e0.voxel_grid, wl, traj, scans)
    e0.close()
    c0, s0, gt = queries[0]
    init = S.perturb_pose(gt, np.random.default_rng(1))
    for G in (0, 2, 4, 8, 16):
        e1 = Engine(default_params(lm_cluster=G, **bench.OVER)); e1.set_submap(mc, ms)
        for _ in range(5): e1.scan2map(c0, s0, init, want_stats=False)
        t0=time.perf_counter()
        for _ in range(50): e1.scan2map(c0, s0, init, want_stats=False)
        t=(time.perf_counter()-t0)/50*1e6
        t0=time.perf_counter()
        for _ in range(20):
            e1.set_submap(mc, ms); e1.scan2map(c0, s0, init, want_stats=False)
        t2=(time.perf_counter()-t0)/20*1e6
        print(wl, 'G', G, 'us/scan %.1f'%t, 'with submap build %.1f'%t2, 'queries', c0.shape[0]+s0.shape[0])
        e1.close()
PY
