timeout 900 python -m pytest tests/test_growth_gpu.py -m gpu -x -q 2>&1 | tail -3
