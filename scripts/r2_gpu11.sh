mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "golden or lockstep_isolated or chained or restatement" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2s_pytest.log
timeout 600 python bench.py --steps 6 --warmup 3 --exact-steps 0 --cpu-budget 0 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'ms/substep', round(d['config']['ms_per_substep'],3), 'its', d['config']['pcg_iterations_timed'])
print('  ', {k: round(v['avg_ms'],4) for k,v in d['kernels'].items()})
print('  ', {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2s_launches.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/r2s_prof.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py launches gpurun_out/r2s_launches.csv gpurun_out/r2s_launches.md; grep -E "p2g|sdf|occ|build_src|tile|total" gpurun_out/r2s_launches.md
