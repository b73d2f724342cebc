# developer batch: the friction tests and the golden-stage test (the constraint's default path)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_moving_solids_gpu.py tests/test_gpu_parity.py -m gpu -q --tb=short -k "friction or stages_against_golden" > gpurun_out/r2j_friction.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2j_friction.log | cut -c1-1800
