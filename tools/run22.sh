timeout 120 python -m pytest tests/test_linear_gpu.py -x -q 2>&1 | tail -15
timeout 120 python tools/quick_lin.py 2>&1 | tail -8
