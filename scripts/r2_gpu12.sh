mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_p2g_scatter|k_p2g_finish|k_sdf_shell|k_p2g_literal" -c 4 -o gpurun_out/r2n_p2g python scripts/profile_step.py sphere256 3 1 > gpurun_out/r2n_prof.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_brief.py gpurun_out/r2n_p2g.ncu-rep > gpurun_out/r2n_brief.txt 2>&1
