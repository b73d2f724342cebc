# developer batch: the moving-solid tests plus the parity tests nearest to the code they touch
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_moving_solids_gpu.py -m gpu -q --tb=short > gpurun_out/r2d_moving.log 2>&1; echo "moving rc=$?"; tail -40 gpurun_out/r2d_moving.log | cut -c1-1500
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "stages_against_golden or static_obstacles or lockstep_chained" > gpurun_out/r2d_near.log 2>&1; echo "near rc=$?"; tail -3 gpurun_out/r2d_near.log | cut -c1-600
