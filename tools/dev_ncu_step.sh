# dev: full ncu capture (with source) of one outer iteration's kernels of the bench step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:k_lm_solve|k_knn5|k_fit|k_transform_keys" -s 8 -c 4 -f -o gpurun_out/step_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log | cut -c1-300
ls -la gpurun_out/
