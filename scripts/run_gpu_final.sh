# round-end style validation on one B200: tests, smoke, bench (both arms), memcheck, ncu launch list
mkdir -p gpurun_out
T=${1:-s}
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "Fluid Engine\|^---" | tail -3
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/profile_step.py dam32 0 2 > gpurun_out/${T}_sanitize.log 2>&1; echo "sanitize rc=$?"; tail -3 gpurun_out/${T}_sanitize.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/${T}_prof.log 2>&1; echo "ncu list rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'], 'scaling_ref', d.get('scaling_ref'))
print('roofline', d['roofline']); print('cpu', d['cpu_baseline']); print('launches', d['gpu_launches'], d['clocks'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'], round(v.get('frac',0),3))
P
