mkdir -p gpurun_out
python bench.py --workload hdl64 --batch 512 --steps 5 --warmup 3 --cpu-sample 64 > gpurun_out/bench_hdl64.json 2> gpurun_out/bench_hdl64.err; tail -2 gpurun_out/bench_hdl64.err
MSFL_BENCH_LM_CLUSTER=4 python bench.py --workload os1-128 --map-scans 50 --batch 128 --distinct 8 --steps 5 --warmup 3 --cpu-sample 16 > gpurun_out/bench_os1.json 2> gpurun_out/bench_os1.err; tail -2 gpurun_out/bench_os1.err
python - <<'PY'
import json
for f in ('hdl64','os1'):
    try:
        d=json.loads(open(f'gpurun_out/bench_{f}.json').read().strip().splitlines()[-1]); r=d['roofline']
        print(f, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r['kernel'][:14], r['frac'], d['config']['queries_per_scan'], d['config']['submap_points'], d['cpu_baseline']['value'], d['pose_err_vs_oracle'])
    except Exception as ex: print(f, 'failed', ex)
PY
