mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2ae_gpu_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2ae_gpu_tests.log | cut -c1-900
