# developer batch: the free-running moving-solid tests with their measured differences printed
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_moving_solids_gpu.py -m gpu -x -q -s --tb=short -k "animated or through_the_api" 2>&1 | grep -E "measured|passed|failed|Error|assert" > gpurun_out/r2l_measured.log; cat gpurun_out/r2l_measured.log | cut -c1-200
