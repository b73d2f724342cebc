mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2ai_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ai_gpu_tests.log | cut -c1-600
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-budget 2 --exact-steps 0 > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ai_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d['e2e'])
print({k:round(v,3) for k,v in d['stage_ms_per_step'].items()})
print({k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})
PY
timeout 900 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:k_ext -c 9 --csv --log-file gpurun_out/r2ai_ext.csv python bench.py --steps 1 --warmup 1 --no-parity-check --cpu-budget 0 --exact-steps 0 > gpurun_out/r2ai.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2ai_ext.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:30]),{})[r[mi]]=r[vi]
for k,v in d.items(): print(k, v)
PY
