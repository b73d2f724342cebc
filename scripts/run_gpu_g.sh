python scripts/e2e_breakdown.py 256 2>&1 | grep -v "Fluid Engine\|^---" | tail -6
