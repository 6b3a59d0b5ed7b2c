mkdir -p gpurun_out
python -m pytest tests/test_async_gpu.py tests/test_scan2map_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu --no-workloads --steps 20 > gpurun_out/r2c14_default.json 2> gpurun_out/r2c14_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c14_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], 'pcl', d['e2e']['pcl_layout_value'])
PY
ncu --set full --clock-control none -k "regex:k_lm_solve|k_knn5|k_fit|k_transform_keys|k_scatter_perm" -s 24 -c 12 -f -o gpurun_out/r2_hdl64_prof \
    python bench.py --workload hdl64 --batch 512 --distinct 8 --steps 2 --warmup 3 --no-cpu --only-device > gpurun_out/r2_ncu_hdl64.log 2>&1
ncu --set full --clock-control none -k "regex:k_lm_solve|k_knn5|k_fit|k_transform_keys|k_scatter_perm" -s 24 -c 12 -f -o gpurun_out/r2_os1_prof \
    python bench.py --workload os1-128 --batch 64 --distinct 4 --steps 2 --warmup 3 --no-cpu --only-device > gpurun_out/r2_ncu_os1.log 2>&1
tail -2 gpurun_out/r2_ncu_os1.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
