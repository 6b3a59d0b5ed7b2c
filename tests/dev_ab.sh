python -m pytest tests/test_features_gpu.py -m gpu -x -q 2>&1 | tail -2
python tests/dev_stgm.py
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/stgm_launches.csv python tests/dev_stgm.py > /dev/null 2>&1
