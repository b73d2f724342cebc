mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/slab_check.py damz128 6 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/j_bench${N}.json 2> gpurun_out/j_bench$N.err; echo "bench$N rc=$?"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/j_bench$N.err | tail -5
python - <<P
import json
d=json.loads(open('gpurun_out/j_bench${N}.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['config']['pcg_iterations_timed'])
print(d['stage_ms_per_step'])
print({k:(round(v['avg_ms'],4),v['launches']) for k,v in d['kernels'].items()})
P
