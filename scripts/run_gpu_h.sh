mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab_multigpu.py -m gpu -x -q --tb=short 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err; echo "bench2 rc=$?"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/h_bench2.err | tail -5
python - <<P
import json
d=json.loads(open('gpurun_out/h_bench2.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], d['config']['pcg_iterations_timed'])
print(d['stage_ms_per_step'])
print({k:(round(v['avg_ms'],4),v['launches']) for k,v in d['kernels'].items()})
P
