timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Fluid Engine\|^---" | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "restatement" 2>&1 | tail -3
