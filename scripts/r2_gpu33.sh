mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "odd_grids" > gpurun_out/r2ap_odd.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2ap_odd.log | cut -c1-500
python scripts/e2e_breakdown.py 256 2>&1 | tail -5
