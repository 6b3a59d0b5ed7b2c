python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --cpu-sample 256 > gpurun_out/bench_dev.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dev.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['synchronous_call_value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()}, d['pose_err_vs_oracle'], d['gpu_launches'])
PY
