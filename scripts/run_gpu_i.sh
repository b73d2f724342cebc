mkdir -p gpurun_out
N=${1:-8}
for lc in ${2:-2}; do
FLIP_MG_GLOBAL_FROM=$lc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/i_bench${N}_lc$lc.json 2> gpurun_out/i_bench$N.err; echo "bench$N lc=$lc rc=$?"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/i_bench$N.err | tail -5
python - <<P
import json
d=json.loads(open('gpurun_out/i_bench${N}_lc$lc.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['config']['pcg_iterations_timed'])
print(d['stage_ms_per_step'])
print({k:(round(v['avg_ms'],4),v['launches']) for k,v in d['kernels'].items()})
P
done
