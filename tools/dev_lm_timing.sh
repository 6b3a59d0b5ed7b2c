MSFL_NVCC_EXTRA=-DMSFL_LM_TIMING python -m msf_loam_b200.build --force > /dev/null
python - <<'PY'
import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from conftest import make_map_case
from msf_loam_b200 import Engine, default_params
case = make_map_case()
q = case["queries"][0]
e = Engine(default_params(early_exit=0, max_num_iterations=5))
e.set_submap(case["map_corner"], case["map_surf"])
for _ in range(2):
    e.scan2map(q["corner"], q["surf"], q["init"], want_stats=False)
e.close()
PY
