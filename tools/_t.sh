python -m pytest tests/test_gpu_parity.py -x -q -k "brick_pair" 2>&1 | tail -12
