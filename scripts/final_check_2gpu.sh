# developer batch: the 2-GPU parity tests at the final revision (gpurun --gpus 2)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_slab_multigpu.py -m gpu -x -q --tb=short > gpurun_out/final_slab_tests.log 2>&1; echo "slab tests rc=$?"; grep -v "^\[rank\|Warning" gpurun_out/final_slab_tests.log | tail -3 | cut -c1-300
