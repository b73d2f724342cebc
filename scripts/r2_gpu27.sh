mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-budget 2 --exact-steps 0 > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2aj_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d.get('parity_ok'), d['e2e']['ms_per_step'])
print({k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_ext -c 4 -o gpurun_out/r2aj_ext -f python bench.py --steps 1 --warmup 1 --no-parity-check --cpu-budget 0 --exact-steps 0 > gpurun_out/r2aj.log 2>&1; echo "rc=$?"
python scripts/ncu_brief.py gpurun_out/r2aj_ext.ncu-rep k_ext
