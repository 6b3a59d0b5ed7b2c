run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['synchronous_call_value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()}, d['clocks'])
    else: print(l.rstrip())
"; }
run MSFL_BENCH_LM_CLUSTER=1
run MSFL_BENCH_LM_CLUSTER=2
run MSFL_BENCH_LM_CLUSTER=4
