B="python bench.py --steps 5 --warmup 3 --equil 2000 --no-secondary --no-cpu --e2e-steps 0 --lammps-steps 0"
OBMD_B200_USH_G=16 ncu --set full --import-source on --clock-control none -k regex:"k_usher_eval" --launch-skip 2005 -c 1 -o gpurun_out/r2_ush16 -f $B > gpurun_out/r2_ush16.log 2>&1; echo rc=$?
