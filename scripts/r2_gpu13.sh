mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2aa_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2aa_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2aa_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'value', d['value'], 'e2e', d['e2e'], 'exact', d.get('exact_mode'))
print('roofline', d['roofline'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'], round(v.get('frac',0),3))
P
