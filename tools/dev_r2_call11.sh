mkdir -p gpurun_out
for c in 0 3 2; do
  MSFL_LM_CTAS_PER_SM=$c python bench.py --no-cpu --no-workloads --only-device --steps 20 > gpurun_out/r2c11_ctas$c.json 2> gpurun_out/r2c11_ctas$c.err
done
for B in 1776 1332; do
  MSFL_LM_CTAS_PER_SM=3 python bench.py --no-cpu --no-workloads --only-device --steps 20 --batch $B > gpurun_out/r2c11_ctas3_b$B.json 2> gpurun_out/r2c11_ctas3_b$B.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c11_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], d['value'], d['ms_per_step'], r['stage_ms_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
