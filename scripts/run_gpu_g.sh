./build/fluidmanager_headless 100 30 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "fluidmanager or abi" 2>&1 | tail -6
