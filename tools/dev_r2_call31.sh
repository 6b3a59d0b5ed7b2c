mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_deskew.py tests/test_ref_matchers.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/dev_two_streams.py 2048 20 2>&1 | tail -12
