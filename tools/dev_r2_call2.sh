# round 2, call 2: GPU tests with the fast plane fit + seeded search, then A/B of the knobs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c2_pytest.log
python bench.py --no-cpu --steps 10 --batch 2048 > gpurun_out/r2c2_default.json 2> gpurun_out/r2c2_default.err
MSFL_SEED_KNN=0 python bench.py --no-cpu --steps 10 --batch 2048 > gpurun_out/r2c2_noseed.json 2> gpurun_out/r2c2_noseed.err
MSFL_RESORT_OUTER=0 python bench.py --no-cpu --steps 10 --batch 2048 > gpurun_out/r2c2_noresort.json 2> gpurun_out/r2c2_noresort.err
python bench.py --steps 10 --batch 2960 --cpu-sample 256 > gpurun_out/r2c2_b2960.json 2> gpurun_out/r2c2_b2960.err
cat gpurun_out/r2c2_pytest.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c2_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d['roofline']
        print(f.split('/')[-1], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()}, d.get('pose_err_vs_oracle'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/r2c2_*.err
