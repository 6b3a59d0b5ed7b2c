mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-workloads --only-device --steps 20 > gpurun_out/r2c16_default.json 2> gpurun_out/r2c16_default.err
tail -2 gpurun_out/r2c16_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c16_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'])
PY
