python -m pytest tests/test_full_size_gpu.py -m gpu -x -q 2>&1 | tail -15
