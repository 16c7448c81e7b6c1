python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --equil 2000 --no-secondary --no-cpu --e2e-steps 0 --lammps-steps 0"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 56000 -c 300 --csv --log-file gpurun_out/r2_launches_steady.csv $B > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"; tail -2 gpurun_out/r2_ncu_launch.log | cut -c1-200
