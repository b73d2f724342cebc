mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/u_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/u_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'scaling_ref', d.get('scaling_ref',{}).get('ms_per_step'))
print(d['stage_ms_per_step'])
P
