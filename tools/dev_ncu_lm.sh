# dev: full ncu capture (with source) of one k_lm_solve launch of the bench step
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_lm_solve -s 6 -c 1 -f -o gpurun_out/lm_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_lm.log 2>&1
tail -3 gpurun_out/ncu_lm.log
ls -la gpurun_out/
