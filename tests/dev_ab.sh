python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in "0 0" "1 0" "1 1"; do set -- $v; echo "== LM_VARIANT=$1 COMPACT=$2"; MSFL_LM_VARIANT=$1 MSFL_COMPACT=$2 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], r['stage_share'])
    else: print(l.rstrip())
"; done
python bench.py --steps 10 --warmup 3 --cpu-sample 256 | tail -c 900
