# developer batch: the 2-GPU parity test and a short slab bench at the final revision (gpurun --gpus 2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_slab_multigpu.py -m gpu -x -q --tb=short > gpurun_out/final_slab_tests.log 2>&1; echo "slab tests rc=$?"; grep -v "^\[rank\|Warning" gpurun_out/final_slab_tests.log | tail -3 | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 3 --warmup 3 --workload dambreak --grid 256 > gpurun_out/final_bench_n2_256.json 2> gpurun_out/final_bench_n2_256.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/final_bench_n2_256.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'parity_ok', d.get('parity_ok'), d['config'].get('workload'))
P
