mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2aq_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r2aq_gpu_tests.log | cut -c1-700
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 0 --exact-steps 0 > gpurun_out/r2aq_bench.json 2> gpurun_out/r2aq_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aq_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
print({k:round(v,3) for k,v in d['stage_ms_per_step'].items()})
print({k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})
PY
python scripts/e2e_breakdown.py 256 2>&1 | tail -5
