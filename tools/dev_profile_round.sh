# dev: the round's profile artefacts (ncu launch list + full-set capture of the step kernels) -> gpurun_out/
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k "regex:k_lm_solve|k_knn5|k_fit|k_transform_keys|k_scatter_perm" -s 20 -c 5 -f -o gpurun_out/step_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_step.log 2>&1
ls -la gpurun_out
