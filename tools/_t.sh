python -m pytest tests/test_gpu_parity.py -x -q -k "bond_harmonic" 2>&1 | tail -15
