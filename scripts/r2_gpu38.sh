mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 0 --exact-steps 0 > gpurun_out/r2au_bench.json 2> gpurun_out/r2au_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2au_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d['roofline'].get('traffic'))
print({k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})
PY
