# round 2, call 4 (2 GPUs): C-ABI broadcast test + the N=2 bench line
mkdir -p gpurun_out
python -m pytest tests/test_bcast_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c4_pytest.log
cat gpurun_out/r2c4_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2c4_n2.json 2> gpurun_out/r2c4_n2.err
tail -5 gpurun_out/r2c4_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c4_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['impl_notes'], d['pose_err_vs_oracle'])
for k,w in (d.get('workloads') or {}).items():
    print('   ',k, w.get('value'), w.get('ms_per_step'), (w.get('e2e') or {}).get('value'), w.get('pose_err_vs_oracle'))
PY
