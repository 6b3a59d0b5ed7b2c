mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2c30_pytest.log 2>&1
tail -5 gpurun_out/r2c30_pytest.log
( time python bench.py ) > gpurun_out/r2c30_bench.json 2> gpurun_out/r2c30_bench.err
tail -c 600 gpurun_out/r2c30_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c30_bench.json').read().strip().splitlines()[0])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
