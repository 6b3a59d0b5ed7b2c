# dev: phase cycles of the fused single-scan kernel (library built with MSFL_NVCC_EXTRA=-DMSFL_FUSED_TIMING and copied to
# msf_loam_b200/libmsfl_ftiming.so)
MSFL_LIB_PATH=$PWD/msf_loam_b200/libmsfl_ftiming.so python - <<'PY'
import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from conftest import make_map_case
from msf_loam_b200 import Engine, default_params
case = make_map_case()
q = case["queries"][0]
for over in ({}, {"early_exit": 0, "max_num_iterations": 5}):
    e = Engine(default_params(lm_cluster=16, **over))
    e.set_submap(case["map_corner"], case["map_surf"])
    print("params", over, flush=True)
    for _ in range(3):
        e.scan2map(q["corner"], q["surf"], q["init"], want_stats=False)
    e.close()
PY
