python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --cpu-sample 256 > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err; tail -3 gpurun_out/bench_dev.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dev.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], r['stage_share'], d['pose_err_vs_oracle'])
PY
