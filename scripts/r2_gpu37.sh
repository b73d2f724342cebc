mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 python bench.py > gpurun_out/r2at_bench.json 2> gpurun_out/r2at_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2at_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2at_launches.csv python bench.py --steps 2 --warmup 1 --no-parity-check --cpu-budget 0 --exact-steps 0 > gpurun_out/r2at_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -c 3000 --csv --log-file gpurun_out/r2at_traffic.csv python bench.py --steps 1 --warmup 1 --no-parity-check --cpu-budget 0 --exact-steps 0 > gpurun_out/r2at_ncu2.log 2>&1; echo "ncu2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2at_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d.get('cpu_baseline'), d.get('exact_mode'))
print(d['roofline'])
print(d['clocks'], d['gpu_launches'])
PY
