mkdir -p gpurun_out
timeout 1200 python bench.py --gpus 1 --workload dambreak --grid 512 --steps 4 --warmup 3 --exact-steps 0 --cpu-budget 0 > gpurun_out/r2as_bench_n1.json 2> gpurun_out/r2as_bench_n1.err; echo "bench n1 rc=$?"; tail -2 gpurun_out/r2as_bench_n1.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2as_bench_n1.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', d['value'], 'workload', d['config'].get('workload'))
print('  stages', {k: round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()})
print('  kernels', {k: round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})
P
