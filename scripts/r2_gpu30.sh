mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab_multigpu.py -m gpu -x -q --tb=short > gpurun_out/r2am_slab_tests.log 2>&1; echo "slab tests rc=$?"; grep -v "^\[rank\|Warning" gpurun_out/r2am_slab_tests.log | tail -12 | cut -c1-700
