mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_objects.py 3 > gpurun_out/r2af_memcheck_objects.log 2>&1; echo "memcheck rc=$?"; grep -E "frame|done|ERROR SUMMARY" gpurun_out/r2af_memcheck_objects.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_objects.py 2 > gpurun_out/r2af_racecheck_objects.log 2>&1; echo "racecheck rc=$?"; grep -E "done|SUMMARY|hazard" gpurun_out/r2af_racecheck_objects.log | sort | uniq -c | tail -8
timeout 900 compute-sanitizer --tool synccheck python scripts/sanitize_objects.py 2 > gpurun_out/r2af_synccheck_objects.log 2>&1; echo "synccheck rc=$?"; grep -E "done|ERROR SUMMARY" gpurun_out/r2af_synccheck_objects.log | tail -3
timeout 900 compute-sanitizer --tool initcheck python scripts/sanitize_objects.py 2 > gpurun_out/r2af_initcheck_objects.log 2>&1; echo "initcheck rc=$?"; grep -E "done|ERROR SUMMARY" gpurun_out/r2af_initcheck_objects.log | tail -3
