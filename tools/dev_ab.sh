# dev A/B: the bench's device-timed leg with the tree's library and with a variant built earlier
# (MSFL_NVCC_EXTRA=... python -m msf_loam_b200.build --force; cp msf_loam_b200/libmsfl.so msf_loam_b200/libmsfl_<name>.so)
mkdir -p gpurun_out
for v in "$@" ""; do
  lib=msf_loam_b200/libmsfl${v:+_$v}.so
  for rep in 1 2; do
    MSFL_LIB_PATH=$PWD/$lib python bench.py --no-cpu --no-workloads --only-device --steps 30 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$lib', d['value'], d['ms_per_step'], [round(v,4) for v in r['stage_ms_per_step'].values()])"
  done
done
