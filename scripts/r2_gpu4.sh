timeout 300 python scripts/probe_vcycle.py 2>&1 | tail -1 | sed "s/^/fused /"
FLIP_PCG_UNFUSED=1 timeout 300 python scripts/probe_vcycle.py 2>&1 | tail -1 | sed "s/^/unfused /"
FLIP_MG_GROUP=32 timeout 300 python scripts/probe_vcycle.py 2>&1 | tail -1 | sed "s/^/fused group32 /"
FLIP_MG_TRACE=1 timeout 300 python scripts/profile_step.py sphere256 2 1 2>&1 | grep "phase ns" | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or solver_modes or chained" 2>&1 | tail -2
