# round 2, call 3: tests + full default bench (workloads record) + fused-fit A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c3_pytest.log
cat gpurun_out/r2c3_pytest.log
( time python bench.py ) > gpurun_out/r2c3_default.json 2> gpurun_out/r2c3_default.err
tail -5 gpurun_out/r2c3_default.err
MSFL_FUSE_FIT=0 python bench.py --no-cpu --no-workloads --steps 20 > gpurun_out/r2c3_nofuse.json 2> gpurun_out/r2c3_nofuse.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2c3_reference.json 2> gpurun_out/r2c3_reference.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f.split('/')[-1], d['value'], d['ms_per_step'], 'e2e', d.get('e2e'), r.get('kernel','')[:12], r.get('avg_launch_ms'), r.get('frac'), r.get('stage_ms_per_step'), d.get('pose_err_vs_oracle'), d.get('cpu_baseline'), d.get('single_scan_latency'))
        for k,w in (d.get('workloads') or {}).items():
            rr=w.get('roofline') or {}
            print('   ',k, w.get('value'), w.get('ms_per_step'), (w.get('e2e') or {}).get('value'), (w.get('e2e') or {}).get('pcl_layout_value'), rr.get('kernel','')[:12], rr.get('frac'), rr.get('whole_step_frac'), w.get('pose_err_vs_oracle'), (w.get('config') or {}).get('queries_per_scan'), (w.get('config') or {}).get('submap_points'))
    except Exception as e:
        print(f, 'ERR', e)
PY
