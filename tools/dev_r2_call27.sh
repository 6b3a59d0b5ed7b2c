mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_features_gpu.py -x -q -m gpu 2>&1 | tail -6
python - <<'PY'
import time, numpy as np, torch
from msf_loam_b200 import Engine, default_params, synth as S
sc = S.make_scene(); traj = S.trajectory(8)
scans = [S.raycast_scan(sc, "vlp16", traj[k], seed=100 + k) for k in range(8)]
B = 256
e = Engine(default_params())
batch = e.prepare_raw_batch([scans[i % 8][0].copy() for i in range(B)], [scans[i % 8][1].copy() for i in range(B)])
import ctypes as C
from msf_loam_b200 import _lib
def run():
    feats = (_lib.Features * B)()
    e._check(e.lib.msfl_extract_features_batch(e.h, C.c_int(B), batch["arr"], None, feats))
for name in ("first (big shape)", "adapted", "adapted"):
    torch.cuda.synchronize(); t0 = time.perf_counter(); run(); torch.cuda.synchronize()
    print(name, "extract batch 256: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
PY
