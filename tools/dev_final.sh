# dev: end-of-round verification + artefacts
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['synchronous_call_value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()}, d['cpu_baseline'], d['pose_err_vs_oracle'], d['gpu_launches'], d['clocks'])
print(open('gpurun_out/bench_reference.json').read()[:600])
PY
