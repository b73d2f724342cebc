mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2c_full_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c_full_gpu_tests.log | cut -c1-400
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2c_smoke.log
