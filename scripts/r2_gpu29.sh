mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab_multigpu.py -m gpu -x -q --tb=short > gpurun_out/r2al_slab_tests.log 2>&1; echo "slab tests rc=$?"; tail -3 gpurun_out/r2al_slab_tests.log | cut -c1-400
bash scripts/r2_gpu_scale.sh 2 r2al
