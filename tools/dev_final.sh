# dev: end-of-round verification + artefacts (full GPU suite, smoke, default bench, reference arm)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('pcl_layout_value'), r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()}, d['cpu_baseline']['value'], d['pose_err_vs_oracle'], d['gpu_launches'], d['clocks'])
print(d['single_scan_latency'])
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('ms_per_step'))
r=json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1]); print(r['value'], r['cpu_baseline']['cores'], r['cpu_baseline']['reference_compiled'].get('value'))
PY
