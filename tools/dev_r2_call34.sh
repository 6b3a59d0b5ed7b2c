mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r2c34_bench.json 2> gpurun_out/r2c34_bench.err
tail -c 800 gpurun_out/r2c34_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c34_bench.json').read().strip().splitlines()[0])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d['workloads']['deskew_branch_batch_vlp16'])[:900])
PY
( time python bench.py --impl reference ) > gpurun_out/r2c34_reference.json 2> gpurun_out/r2c34_reference.err
tail -3 gpurun_out/r2c34_reference.err; cut -c1-300 gpurun_out/r2c34_reference.json
python -c "
import json; d=json.loads(open('gpurun_out/r2c34_reference.json').read().strip().splitlines()[0]); print(d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['reference_compiled'])"
