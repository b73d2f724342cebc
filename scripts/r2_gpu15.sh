mkdir -p gpurun_out
timeout 600 python scripts/probe_p2g_diff.py sphere256 4 > gpurun_out/r2v_diff.log 2>&1; tail -12 gpurun_out/r2v_diff.log
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2v_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2v_launches.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/r2v_prof.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r2v_launches.csv gpurun_out/r2v_launches.md; grep -E "p2g|sdf|occ|build_src|tile|total" gpurun_out/r2v_launches.md
