mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2e_pytest.log
for mode in scatter gather; do
  if [ $mode = gather ]; then export FLIP_P2G_GATHER=1; else unset FLIP_P2G_GATHER; fi
  timeout 600 python bench.py --steps 6 --warmup 3 --exact-steps 0 --cpu-budget 0 > gpurun_out/r2e_bench_$mode.json 2> gpurun_out/r2e_bench_$mode.err; echo "bench $mode rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/r2e_bench_$mode.json').read().strip().splitlines()[-1])
print('$mode ms/step', round(d['ms_per_step'],3), 'ms/substep', round(d['config']['ms_per_substep'],3), 'its', d['config']['pcg_iterations_timed'])
print('  ', {k: round(v['avg_ms'],4) for k,v in d['kernels'].items()})
print('  ', {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
P
done
unset FLIP_P2G_GATHER
FLIP_MG_TRACE=1 timeout 300 python scripts/profile_step.py sphere256 2 1 2>&1 | grep "phase ns" | tail -1
