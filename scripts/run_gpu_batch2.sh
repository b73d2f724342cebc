FLIP_MG_TRACE=1 python scripts/profile_step.py sphere256 2 1 2>&1 | tail -3
echo "--- no L2 pin"
FLIP_MG_NO_L2PIN=1 FLIP_MG_TRACE=1 python scripts/profile_step.py sphere256 2 1 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-scaling-ref --cpu-steps 1 > gpurun_out/f_bench.json 2>gpurun_out/f_bench.err; tail -2 gpurun_out/f_bench.err
