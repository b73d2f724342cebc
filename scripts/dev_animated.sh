# developer batch: the animated-mesh test, twice (run-to-run spread of the splash)
mkdir -p gpurun_out
for r in 1 2; do timeout 200 python -m pytest tests/test_moving_solids_gpu.py -m gpu -q --tb=short -k "animated" > gpurun_out/r2h_moving_$r.log 2>&1; echo "moving run $r rc=$?"; tail -12 gpurun_out/r2h_moving_$r.log | cut -c1-1500; done
