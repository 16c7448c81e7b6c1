python -m pytest tests/test_lammps_plugin.py -x -q -k argument_errors 2>&1 | tail -3
OBMD_B200_DEBUG_FIX=2090 python bench.py --steps 100 --warmup 5 --equil 2000 --no-secondary --no-cpu --e2e-steps 0 --lammps-steps 0 2>&1 | grep -E "fix round|  trial" | head -400 > gpurun_out/r2_fix_debug.txt; wc -l gpurun_out/r2_fix_debug.txt
