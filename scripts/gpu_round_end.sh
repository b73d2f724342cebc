# developer batch at the end of round 2: the whole GPU suite, smoke(), sanitizers over the moving-solid paths, the short bench
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/r2g_full_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2g_full_gpu_tests.log | cut -c1-400
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2g_smoke.log
timeout 40 python scripts/sanitize_moving.py 3 > gpurun_out/r2g_moving_plain.log 2>&1; echo "plain rc=$?"; tail -7 gpurun_out/r2g_moving_plain.log | cut -c1-300
for tool in memcheck racecheck initcheck; do
  timeout 150 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_moving.py 2 > gpurun_out/r2g_moving_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "SUMMARY|hazard" gpurun_out/r2g_moving_$tool.log | tail -3
done
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-budget 0 --exact-steps 0 > gpurun_out/r2g_bench_short.json 2> gpurun_out/r2g_bench_short.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g_bench_short.json').read().strip().split('\n')[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
print({k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})
PY
