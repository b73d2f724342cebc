mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "odd_grids" > gpurun_out/r2ar_odd.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2ar_odd.log | cut -c1-3000
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 0 --exact-steps 0 > gpurun_out/r2ar_bench.json 2> gpurun_out/r2ar_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ar_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d['e2e'])
PY
