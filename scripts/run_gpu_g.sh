mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "stress" 2>&1 | tail -15
timeout 600 python scripts/pressure_stress.py 256 > gpurun_out/g_stress.jsonl 2> gpurun_out/g_stress.err; echo "stress rc=$?"; tail -3 gpurun_out/g_stress.err; cat gpurun_out/g_stress.jsonl
