mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short -k "stage or lockstep or extrap or frame or default" > gpurun_out/r2ak_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ak_gpu_tests.log | cut -c1-600
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-budget 2 --exact-steps 0 > gpurun_out/r2ak_bench.json 2> gpurun_out/r2ak_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ak_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d.get('parity_ok'), d['e2e']['ms_per_step'])
print({k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})
PY
timeout 900 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:k_ext -c 9 --csv --log-file gpurun_out/r2ak_ext.csv python bench.py --steps 1 --warmup 1 --no-parity-check --cpu-budget 0 --exact-steps 0 > gpurun_out/r2ak.log 2>&1; echo "rc=$?"
grep -o '"k_ext[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","gpu__time_duration.sum","[^"]*","[^"]*"' gpurun_out/r2ak_ext.csv | awk -F'","' '{print $1, $NF}' | head -12
