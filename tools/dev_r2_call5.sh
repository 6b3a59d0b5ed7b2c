# round 2, call 5: GPU tests after the atan2 / adapter changes + ncu launch list and full capture of the step kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c5_pytest.log
cat gpurun_out/r2c5_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --only-device > gpurun_out/r2_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k "regex:k_lm_solve|k_knn5|k_fit|k_transform_keys|k_scatter_perm" -s 24 -c 6 -f -o gpurun_out/r2_step_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu --only-device > gpurun_out/r2_ncu_step.log 2>&1
ls -la gpurun_out | tail -5
