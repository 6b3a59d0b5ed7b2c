# dev: the 8-GPU bench line with the final code (the driver's SCALE run does N = 1, 2, 4, 8 itself)
mkdir -p gpurun_out
N=${1:-8}
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 400 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
d=json.loads([l for l in open(f'gpurun_out/bench_n{N}.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('h2d_GBps_per_rank'), d['e2e'].get('pcl_layout_value'))
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
