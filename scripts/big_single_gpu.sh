# developer batch: a large dam break on one GPU (GRID=640 by default; 768^3 = 454 M nodes, 432 M particles), memory and time
mkdir -p gpurun_out
free -g | head -2
avail=$(free -g | awk '/Mem:/ {print $7}')
if [ "$avail" -lt 96 ]; then echo "not enough host memory ($avail GB): skipped"; exit 0; fi
nvidia-smi --query-gpu=memory.used --format=csv,noheader,nounits -l 1 > gpurun_out/big${GRID:-640}_mem.txt &
SMI=$!
timeout 1500 python bench.py --workload dambreak --grid ${GRID:-640} --steps 2 --warmup 1 --cpu-budget 0 --exact-steps 0 --no-parity-check > gpurun_out/big${GRID:-640}.json 2> gpurun_out/big${GRID:-640}.err; echo "rc=$?"; tail -3 gpurun_out/big${GRID:-640}.err | cut -c1-300
kill $SMI; echo "peak device memory MiB: $(sort -n gpurun_out/big${GRID:-640}_mem.txt | tail -1) of $(nvidia-smi --query-gpu=memory.total --format=csv,noheader)"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/big' + __import__('os').environ.get('GRID', '640') + '.json').read().strip().split('\n')[-1])
    print(d['config']['workload'], d['config']['particles'], 'ms/frame', round(d['ms_per_step'],2), 'value', d['value'], 'rows', d['config']['pressure_rows'])
    print({k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0.05})
except Exception as e:
    print('no line', e)
PY
