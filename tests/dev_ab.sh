python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --cpu-sample 128 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['synchronous_call_value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()}, d['pose_err_vs_oracle']['max_trans_m'])
    else: print(l.rstrip())
"; }
run MSFL_X=4
MSFL_NVCC_EXTRA=-DFIT_MINB=5 python -m msf_loam_b200.build --force
run MSFL_X=5
