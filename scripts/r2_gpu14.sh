mkdir -p gpurun_out
timeout 600 python scripts/probe_p2g_diff.py sphere256 4 > gpurun_out/r2u_diff.log 2>&1; tail -45 gpurun_out/r2u_diff.log
