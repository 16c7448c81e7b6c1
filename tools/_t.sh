OBMD_B200_DEBUG_FIX=2080 python bench.py --steps 100 --warmup 5 --equil 2000 --no-secondary --no-cpu --e2e-steps 0 --lammps-steps 0 2>&1 | grep "fix round" | head -40
