# round-2 captures (one GPU): $1 = hot | stream.  hot: the three hot kernels at steady state (2000 equilibration
# steps).  stream: the streaming / bookkeeping kernels and the launch list of five steady-state steps.
# (two calls: gpurun brings back at most 64 MiB)   Usage: gpurun -- 'bash tools/ncu_r2_capture.sh hot'
B="python bench.py --steps 5 --warmup 3 --equil 2000 --no-secondary --no-cpu --e2e-steps 2 --lammps-steps 0"
if [ "$1" = "hot" ]; then
ncu --set full --import-source on --clock-control none -k regex:"k_pair_brick|k_usher_eval|k_usher_fix" --launch-skip 6030 -c 5 -o gpurun_out/r2_hot -f $B > gpurun_out/r2_ncu_hot.log 2>&1; echo "ncu hot rc=$?"
else
ncu --set full --clock-control none -k regex:"k_weight|k_apply|k_gather|k_nve|k_bin|k_count_regions|k_cell_order|k_perm|k_pbc|k_delete_fixup|k_pair_add_new|k_pair_new_atoms" --launch-skip 24100 -c 14 -o gpurun_out/r2_stream -f $B > gpurun_out/r2_ncu_stream.log 2>&1; echo "ncu stream rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60000 -c 400 --csv --log-file gpurun_out/r2_launches_steady.csv $B > gpurun_out/r2_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
fi
