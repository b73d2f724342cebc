mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2ab_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --exact-steps 0 --cpu-budget 0 > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2ab_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2ab_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'value', d['value'], 'its', d['config']['pcg_iterations_timed'], 'substeps', d['config']['substeps_timed'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'], round(v.get('frac',0),3))
print({k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
P
FLIP_MG_TRACE=1 timeout 300 python scripts/profile_step.py sphere256 2 1 > gpurun_out/r2ab_trace.log 2>&1; grep "phase ns" gpurun_out/r2ab_trace.log | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2ab_launches.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/r2ab_prof.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r2ab_launches.csv gpurun_out/r2ab_launches.md; head -30 gpurun_out/r2ab_launches.md
