mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c_pytest.log
for g in 4 8 16 32 148; do FLIP_MG_GROUP=$g timeout 300 python scripts/probe_vcycle.py 2>&1 | tail -1 | sed "s/^/group=$g /"; done
FLIP_PCG_UNFUSED=1 timeout 300 python scripts/probe_vcycle.py 2>&1 | tail -1 | sed "s/^/unfused /"
FLIP_MG_TRACE=1 timeout 300 python scripts/profile_step.py sphere256 2 1 2>&1 | grep "phase ns" | tail -2
