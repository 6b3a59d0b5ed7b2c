python -m pytest tests/test_features_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -3
python tests/dev_odo.py vlp16 20
python tests/dev_odo.py hdl64 10
