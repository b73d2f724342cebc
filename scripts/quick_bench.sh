# developer batch: a few parity tests that exercise the PCG, then the short bench (no CPU arm)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "stages_against_golden or lockstep_chained or solver_modes or pressure" > gpurun_out/quick_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/quick_tests.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 0 --exact-steps 0 > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/quick_bench.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], 'pcg launches', d['kernels']['pcg_iter']['launches'])
print({k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})
PY
