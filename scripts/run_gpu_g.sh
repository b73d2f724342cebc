timeout 1200 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -4
python scripts/probe_vcycle.py 2>&1 | tail -1
