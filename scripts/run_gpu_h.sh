mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/h_bench2_long.json 2> gpurun_out/h_bench2.err; echo "bench2 rc=$?"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/h_bench2.err | tail -5
python - <<P
import json
d=json.loads(open('gpurun_out/h_bench2_long.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'], 'substeps', d['config']['substeps_timed'], 'pcg', d['config']['pcg_iterations_timed'], 'particles', d['config']['particles'])
P
