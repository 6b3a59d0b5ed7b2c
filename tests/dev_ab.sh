run() { echo "== $*"; python bench.py --steps 10 --warmup 3 --cpu-sample 64 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()}, d['pose_err_vs_oracle']['max_trans_m'])
    else: print(l.rstrip())
"; }
MSFL_NVCC_EXTRA=-DMSFL_LM_THREADS=64 python -m msf_loam_b200.build --force > /dev/null
run threads64
python -m pytest tests/test_scan2map_gpu.py -m gpu -x -q 2>&1 | tail -2
MSFL_NVCC_EXTRA=-DMSFL_LM_THREADS=256 python -m msf_loam_b200.build --force > /dev/null
run threads256
