python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()})
    else: print(l.rstrip())
"; }
run MSFL_COUNT_SORT=0
run MSFL_COUNT_SORT=1
python bench.py --steps 10 --warmup 3 --cpu-sample 256 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['pose_err_vs_oracle'])"
