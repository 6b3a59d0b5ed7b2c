mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2c13_pytest.log
cat gpurun_out/r2c13_pytest.log
python bench.py --no-cpu --no-workloads --steps 20 > gpurun_out/r2c13_default.json 2> gpurun_out/r2c13_default.err
tail -3 gpurun_out/r2c13_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c13_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
PY
