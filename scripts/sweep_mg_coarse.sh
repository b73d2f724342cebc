mkdir -p gpurun_out
for cfg in "148 148" "148 32" "148 16" "148 8" "111 16" "74 16" "74 74"; do
set -- $cfg
FLIP_MG_COARSE_BLOCKS=$1 FLIP_MG_GROUP=$2 timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 0 --exact-steps 0 --no-parity-check > gpurun_out/sweep.json 2> gpurun_out/sweep.err
python - <<PY
import json
d=json.loads(open('gpurun_out/sweep.json').read().strip().split('\n')[-1])
print("blocks $1 group $2:", round(d['ms_per_step'],3), 'precond', round(d['kernels']['precond']['avg_ms'],4), 'pcg_iter', round(d['kernels']['pcg_iter']['avg_ms'],4))
PY
done
