mkdir -p gpurun_out
for a in 0 0.5 1.0; do
FLIP_WARM_EXTRAPOLATE=$a timeout 600 python bench.py --steps 10 --warmup 3 --cpu-budget 0 --exact-steps 0 > gpurun_out/r2ao_bench_$a.json 2> gpurun_out/r2ao_bench_$a.err; echo "bench a=$a rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2ao_bench_$a.json').read().strip().split('\n')[-1])
print('a=$a', d['ms_per_step'], d['value'], 'pcg launches', d['kernels']['pcg_iter']['launches'], 'pressure', round(d['stage_ms_per_step']['pressure'],3), d['config'].get('pcg_iterations'))
PY
done
