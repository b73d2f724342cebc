mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2ag_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ag_gpu_tests.log | cut -c1-600
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-budget 2 --exact-steps 0 > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ag_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d['e2e'])
print({k:round(v,3) for k,v in d['stage_ms_per_step'].items()})
print({k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})
PY
