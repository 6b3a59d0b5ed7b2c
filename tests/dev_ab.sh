python -m pytest tests/test_features_gpu.py tests/test_golden.py tests/test_io.py -m gpu -x -q 2>&1 | tail -2
python tests/dev_odo.py vlp16 20 | tail -3
python tests/dev_odo.py hdl64 10 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/odo_launches.csv python tests/dev_odo.py vlp16 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/odo_launches_hdl.csv python tests/dev_odo.py hdl64 1 > /dev/null 2>&1
