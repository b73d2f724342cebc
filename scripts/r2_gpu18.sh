mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "seeding or surface_reconstruction or headless" > gpurun_out/r2z_new.log 2>&1; echo "new tests rc=$?"; tail -30 gpurun_out/r2z_new.log | cut -c1-400
