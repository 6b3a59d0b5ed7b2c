python -m pytest tests/test_scan2map_gpu.py tests/test_golden.py tests/test_deskew.py tests/test_features_gpu.py tests/test_async_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --cpu-sample 256 > gpurun_out/bench_dev.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dev.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['synchronous_call_value'], r['kernel'][:12], r['avg_launch_ms'], r['frac'], {k[:10]:v for k,v in r['stage_share'].items()}, d['pose_err_vs_oracle'], d['gpu_launches'])
PY
python tests/dev_latency.py 2>&1 | grep -E "lm_cluster=1|lm_cluster=8"
bash tests/dev_lm_timing.sh 2>&1 | grep "LM timing" | tail -2
python -m msf_loam_b200.build --force > /dev/null
