mkdir -p gpurun_out
for N in 8 4; do
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 ) > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  tail -4 gpurun_out/r2_scale_n$N.err
done
python - <<'PY'
import json
for N in (8,4):
    try:
        d=json.loads(open(f'gpurun_out/r2_scale_n{N}.json').read().strip().splitlines()[-1])
        print(N, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_GBps_per_rank'], d['e2e']['pcl_layout_value'], d['host'], d['pose_err_vs_oracle'])
        for k,w in (d.get('workloads') or {}).items():
            print('   ',k, w.get('value'), w.get('ms_per_step'), (w.get('e2e') or {}).get('value'))
    except Exception as e: print(N,'ERR',e)
PY
