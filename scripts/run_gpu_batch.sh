mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/profile_step.py dam32 0 2 > gpurun_out/d_sanitize.log 2>&1; echo "sanitize rc=$?"
tail -5 gpurun_out/d_sanitize.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/d_pytest.log
timeout 600 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
cat gpurun_out/d_bench.json | cut -c1-6000
tail -5 gpurun_out/d_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/d_launches.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/d_prof.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/d_prof.log
