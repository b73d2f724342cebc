mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "surface or headless" > gpurun_out/r2ax_surface.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2ax_surface.log | cut -c1-400
ISOMESH_REF=0 timeout 600 python tests/bench_isomesh.py default30:2 dam64:2 dam128:2 spheredrop256:1 2>&1 | grep "^{" | cut -c1-330
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_objects.py 3 2>&1 | grep -E "frame|done|ERROR SUMMARY" | tail -5
