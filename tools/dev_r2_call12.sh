mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2c12_pytest.log
cat gpurun_out/r2c12_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -4 gpurun_out/r2_bench_n1.err
( time python bench.py --impl reference --steps 10 --warmup 2 ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
tail -3 gpurun_out/r2_bench_reference.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['pcl_layout_value'], d['clocks'], d['single_scan_latency'])
print('same config:', d['config']==r['config'], 'reference', r['value'], r['cpu_baseline']['cores'])
for k,w in (d.get('workloads') or {}).items():
    print('   ',k, w.get('value'), w.get('ms_per_step'), (w.get('e2e') or {}).get('value'), w.get('pose_err_vs_oracle'))
PY
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print('with_odometry', json.dumps(d['workloads']['chain_raw_to_pose_vlp16'].get('with_odometry'))[:900])
PY
