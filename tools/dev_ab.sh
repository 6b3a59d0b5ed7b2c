python -m pytest tests/test_c_abi_program.py -m gpu -x -q -s 2>&1 | grep -E "async|passed|failed|pose t" | head -5
run() { echo "== $*"; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], 'e2e', d['e2e']['value'], 'sync', d['e2e']['synchronous_call_value'])
    else: print(l.rstrip())
"; }
run MSFL_X=geo4
run MSFL_CHUNK_SCHED=eq4
run MSFL_CHUNK_SCHED=3
run MSFL_CHUNK_SCHED=2
run MSFL_CHUNK_SCHED=b
