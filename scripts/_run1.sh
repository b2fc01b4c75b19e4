timeout 300 python -m pytest tests/test_extract_gpu.py -m gpu -x -q 2>&1 | tail -3
