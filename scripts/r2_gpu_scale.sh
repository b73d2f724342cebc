# usage: r2_gpu_scale.sh N tag   (under gpurun --gpus N)
N=$1; TAG=$2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
SLAB_CHECK_ORACLE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/slab_check.py damz128 6 > gpurun_out/${TAG}_slab_check_n$N.log 2>&1; echo "slab_check rc=$?"; grep -E "SLAB_CHECK|'frame': 5" gpurun_out/${TAG}_slab_check_n$N.log | cut -c1-400
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_n$N.err
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
    print('N', d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', d['value'], 'parity_ok', d.get('parity_ok'), 'workload', d['config'].get('workload'))
    print('  stages', {k: round(v,3) for k,v in d.get('stage_ms_per_step',{}).items()})
    print('  kernels', {k: round(v['avg_ms'],4) for k,v in d.get('kernels',{}).items()})
except Exception as e:
    print('no bench line', e)
P
