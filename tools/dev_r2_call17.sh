python -m pytest tests/test_c_abi_program.py -m gpu -x -q 2>&1 | tail -40
