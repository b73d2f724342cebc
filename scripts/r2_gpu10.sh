mkdir -p gpurun_out
FLIP_MG_TRACE=1 timeout 300 python scripts/profile_step.py sphere256 2 1 > gpurun_out/r2x_trace.log 2>&1; grep "phase ns" gpurun_out/r2x_trace.log | tail -2
timeout 300 python scripts/probe_vcycle.py > gpurun_out/r2x_vcycle.log 2>&1; tail -1 gpurun_out/r2x_vcycle.log
