mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/r2c38_bench_n2.json 2> gpurun_out/r2c38_bench_n2.err
tail -c 600 gpurun_out/r2c38_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c38_bench_n2.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('h2d_GBps_per_rank'))
for k,v in d['workloads'].items(): print(k, v.get('value'), v.get('ms_per_step'))
PY
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 ) 2>&1 | tail -c 700
