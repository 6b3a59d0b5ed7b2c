# round 2, call 1: GPU tests with the new LM sweep, then A/B of the step against the round-1 library
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c1_pytest.log
for B in 2048 2368 2960; do
  python bench.py --no-cpu --steps 10 --batch $B > gpurun_out/r2c1_new_b$B.json 2> gpurun_out/r2c1_new_b$B.err
done
MSFL_LIB_PATH=$PWD/msf_loam_b200/libmsfl_r1.so python bench.py --no-cpu --steps 10 --batch 2048 > gpurun_out/r2c1_r1_b2048.json 2> gpurun_out/r2c1_r1_b2048.err
MSFL_LIB_PATH=$PWD/msf_loam_b200/libmsfl_r1.so python bench.py --no-cpu --steps 10 --batch 2960 > gpurun_out/r2c1_r1_b2960.json 2> gpurun_out/r2c1_r1_b2960.err
cat gpurun_out/r2c1_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c1_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], r['stage_share'])
    except Exception as e:
        print(f, 'ERR', e)
PY
