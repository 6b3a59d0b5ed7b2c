mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2c10_pytest.log
cat gpurun_out/r2c10_pytest.log
( time python bench.py ) > gpurun_out/r2c10_default.json 2> gpurun_out/r2c10_default.err
tail -4 gpurun_out/r2c10_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c10_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], (d['e2e'] or {}).get('value'), d['single_scan_latency'])
for k,w in (d.get('workloads') or {}).items():
    print('   ',k, w.get('value'), w.get('ms_per_step'), (w.get('e2e') or {}).get('value'), w.get('pose_err_vs_oracle'), w.get('queries_per_scan'))
PY
