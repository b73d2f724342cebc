mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/g_pytest.log
python bench.py --no-scaling-ref --cpu-steps 1 > gpurun_out/g_bench.json 2>gpurun_out/g_bench.err; tail -2 gpurun_out/g_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'], round(v.get('frac',0),3))
P
