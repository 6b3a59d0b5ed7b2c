python -m pytest tests/test_scan2map_gpu.py tests/test_full_size_gpu.py tests/test_edge_cases_gpu.py tests/test_async_gpu.py -m gpu -x -q 2>&1 | tail -2
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()})
    else: print(l.rstrip())
"; }
run MSFL_X=1
run MSFL_KNN_MINB12=1
