mkdir -p gpurun_out
for m in 10 12 14; do
  MSFL_LIB_PATH=$PWD/msf_loam_b200/libmsfl_minb$m.so python bench.py --no-cpu --no-workloads --only-device --steps 20 > gpurun_out/r2c18_minb$m.json 2> gpurun_out/r2c18_minb$m.err
done
python bench.py --no-cpu --no-workloads --only-device --steps 20 > gpurun_out/r2c18_minb8.json 2> gpurun_out/r2c18_minb8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c18_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], d['value'], d['ms_per_step'], list(r['stage_ms_per_step'].values()))
    except Exception as e: print(f,'ERR',e)
PY
