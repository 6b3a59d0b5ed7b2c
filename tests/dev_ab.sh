run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['synchronous_call_value'], r['avg_launch_ms'], d['clocks'])
    else: print(l.rstrip())
"; }
run MSFL_X=1
run MSFL_BENCH_NO_SAMPLER=1
run MSFL_X=1
run MSFL_BENCH_NO_SAMPLER=1
