mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2c9_pytest.log
cat gpurun_out/r2c9_pytest.log
python - <<'PY'
import sys, os, struct, subprocess, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import make_map_case
from msf_loam_b200 import synth as S
for sensor, scene, seed in (("vlp16","room40",100),("hdl64","room80",200)):
    case = make_map_case(sensor, scene, 5, seed); q = case["queries"][0]
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
    print(sensor, subprocess.run(["/tmp/adrv","/tmp/case.bin","bench"],capture_output=True,text=True).stdout.strip().splitlines()[-1])
PY
