mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/g_pytest.log
