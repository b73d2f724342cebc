mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/g_launches.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/g_prof.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/g_prof.log
python bench.py --no-scaling-ref --cpu-steps 1 > gpurun_out/g_bench2.json 2>gpurun_out/g_bench2.err; tail -2 gpurun_out/g_bench2.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/g_bench2.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'], round(v.get('frac',0),3))
P
