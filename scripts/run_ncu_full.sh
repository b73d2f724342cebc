mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"k_sdf_p2g|k_sdf_far|k_g2p_fast|k_advance_fast|k_gather|k_classify|k_cell_finalize|k_ext_claim|k_ext_init|k_build_system|k_apply_pressure|k_seg_flag" -c 14 \
  -f -o gpurun_out/e_particles python scripts/profile_step.py sphere256 3 1 > gpurun_out/e_particles.log 2>&1; echo "ncu A rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"k_pcg_spmv|k_pcg_update|k_pcg_direction|k_mg0_sweep|k_mg0_restrict_list" --launch-skip 14 -c 7 \
  -f -o gpurun_out/e_pcg python scripts/profile_step.py sphere256 3 1 > gpurun_out/e_pcg.log 2>&1; echo "ncu B rc=$?"
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/e_particles.log gpurun_out/e_pcg.log
