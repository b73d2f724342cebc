mkdir -p gpurun_out
timeout 600 python scripts/probe_p2g_diff.py sphere256 4 > gpurun_out/r2w_diff.log 2>&1; tail -9 gpurun_out/r2w_diff.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2w_launches.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/r2w_prof.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r2w_launches.csv gpurun_out/r2w_launches.md; grep -E "p2g|sdf|occ|build_src|tile|total" gpurun_out/r2w_launches.md
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_p2g_scatter|k_sdf_shell" -c 2 -o gpurun_out/r2w_p2g python scripts/profile_step.py sphere256 3 1 > gpurun_out/r2w_prof2.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_brief.py gpurun_out/r2w_p2g.ncu-rep > gpurun_out/r2w_brief.txt 2>&1
