mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:k_ext -c 24 --csv --log-file gpurun_out/r2ah_ext.csv python bench.py --steps 1 --warmup 1 --no-parity-check --cpu-budget 0 --exact-steps 0 > gpurun_out/r2ah.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2ah_ext.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:30]),{})[r[mi]]=r[vi]
for k,v in d.items(): print(k, v)
PY
