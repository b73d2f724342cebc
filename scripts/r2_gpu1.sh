# round 2, GPU call 1: the new headline-size lock-step tests, the new bench contract (both arms, short)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "dambreak128_developed or spheredrop256_headline" > gpurun_out/r2a_pytest_large.log 2>&1; echo "pytest-large rc=$?"
tail -5 gpurun_out/r2a_pytest_large.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2a_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --ref-budget 60 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"; tail -2 gpurun_out/r2a_bench_ref.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'], 'exact', d.get('exact_mode'))
print('roofline', d['roofline']); print('cpu', d.get('cpu_baseline')); print('launches', d['gpu_launches'], d['clocks'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'], round(v.get('frac',0),3))
print(open('gpurun_out/r2a_bench_ref.json').read()[:1500])
P
