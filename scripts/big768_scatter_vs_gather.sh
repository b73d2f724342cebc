mkdir -p gpurun_out
for mode in 0 1; do
if [ $mode = 1 ]; then export FLIP_P2G_GATHER=1; fi
timeout 1500 python bench.py --workload dambreak --grid 768 --steps 1 --warmup 1 --cpu-budget 0 --exact-steps 0 --no-parity-check > gpurun_out/big768_m$mode.json 2> gpurun_out/big768_m$mode.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/big768_m$mode.json').read().strip().split('\n')[-1])
print('gather' if $mode else 'scatter', 'rows', d['config']['pressure_rows'], 'pcg', d['config']['pcg_iterations_timed'], 'substeps', d['config']['substeps_timed'], 'particles', d['config']['particles'], 'ms', round(d['ms_per_step'],1), 'sdf_p2g', round(d['kernels']['sdf_p2g']['avg_ms'],2))
PY
done
