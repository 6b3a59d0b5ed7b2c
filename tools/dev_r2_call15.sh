mkdir -p gpurun_out
python -m pytest tests/test_async_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-workloads --steps 20 > gpurun_out/r2c15_default.json 2> gpurun_out/r2c15_default.err
tail -2 gpurun_out/r2c15_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c15_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], 'f4', d['e2e']['float4_value'], 'pcl', d['e2e']['pcl_layout_value'], 'sync', d['e2e']['synchronous_call_value'])
PY
