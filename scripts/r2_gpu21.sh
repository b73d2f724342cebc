mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "inflow or seeding or headless" > gpurun_out/r2ad_new.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2ad_new.log | cut -c1-900
