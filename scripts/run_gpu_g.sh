python scripts/probe_vcycle.py 2>&1 | tail -1
FLIP_MG_TRACE=1 python scripts/profile_step.py sphere256 2 1 2>&1 | grep "phase ns" | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -3
