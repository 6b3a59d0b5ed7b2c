# final round-2 profile artefacts: ncu launch list + full-set captures of the step kernels for the three bench shapes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only-device > gpurun_out/r2_launches_bench.log 2>&1
K="regex:k_lm_solve|k_knn5|k_fit|k_transform_keys|k_scatter_perm"
ncu --set full --clock-control none --import-source on -k "$K" -s 20 -c 5 -f -o gpurun_out/r2_step_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --only-device > gpurun_out/r2_ncu_step.log 2>&1
ncu --set full --clock-control none -k "$K" -s 20 -c 10 -f -o gpurun_out/r2_hdl64_prof \
    python bench.py --workload hdl64 --batch 512 --distinct 8 --steps 2 --warmup 3 --no-cpu --only-device > gpurun_out/r2_ncu_hdl64.log 2>&1
ncu --set full --clock-control none -k "$K" -s 20 -c 10 -f -o gpurun_out/r2_os1_prof \
    python bench.py --workload os1-128 --batch 64 --distinct 4 --steps 2 --warmup 3 --no-cpu --only-device > gpurun_out/r2_ncu_os1.log 2>&1
ncu --set full --clock-control none -k "regex:k_scan2map_fused" -s 6 -c 1 -f -o gpurun_out/r2_fused_prof \
    python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import make_map_case
from msf_loam_b200 import Engine, default_params
c=make_map_case(); q=c['queries'][0]
e=Engine(default_params(lm_cluster=16)); e.set_submap(c['map_corner'],c['map_surf'])
for _ in range(8): e.scan2map(q['corner'],q['surf'],q['init'],want_stats=False)
" > gpurun_out/r2_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
