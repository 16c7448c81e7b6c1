B="python bench.py --steps 5 --warmup 3 --equil 2000 --no-secondary --no-cpu --e2e-steps 2 --lammps-steps 0"
ncu --set full --import-source on --clock-control none -k regex:"k_pair_row|k_usher_eval|k_usher_fix" --launch-skip 6030 -c 6 -o gpurun_out/r2_hot -f $B > gpurun_out/r2_ncu_hot.log 2>&1; echo "ncu hot rc=$?"
ncu --set full --clock-control none -k regex:"k_weight|k_apply|k_gather|k_nve|k_bin|k_count_regions|k_cell_order|k_perm|k_pbc|k_delete_fixup|k_pair_add_new|k_pair_new_atoms" --launch-skip 24100 -c 24 -o gpurun_out/r2_stream -f $B > gpurun_out/r2_ncu_stream.log 2>&1; echo "ncu stream rc=$?"
ls -la gpurun_out
