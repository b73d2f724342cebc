mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "golden or lockstep_isolated or pressure" > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ac_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --exact-steps 0 --cpu-budget 0 > gpurun_out/r2ac_bench.json 2> gpurun_out/r2ac_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2ac_bench.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2ac_bench.json').read().strip().splitlines()[-1])
print('ms/step', round(d['ms_per_step'],3), 'value', d['value'], 'its', d['config']['pcg_iterations_timed'], 'substeps', d['config']['substeps_timed'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'], round(v.get('frac',0),3))
print({k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"k_mg_invert_small|k_mg_coarse|k_p2g_finish" -c 6 --csv --log-file gpurun_out/r2ac_l.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/r2ac_prof.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2ac_l.csv gpurun_out/r2ac_l.md; grep "flip::" gpurun_out/r2ac_l.md
