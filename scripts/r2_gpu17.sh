mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 900 python -m pytest tests/test_slab_multigpu.py -m gpu -x -q --tb=short > gpurun_out/r2y_slab_pytest.log 2>&1; echo "slab pytest rc=$?"; tail -5 gpurun_out/r2y_slab_pytest.log
for peer in 1 0; do
FLIP_PEER=$peer SLAB_CHECK_ORACLE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/slab_check.py damz64 8 > gpurun_out/r2y_slab_check_n2_peer$peer.log 2>&1; echo "slab_check peer=$peer rc=$?"; grep -E "SLAB_CHECK|frame" gpurun_out/r2y_slab_check_n2_peer$peer.log | tail -4
done
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_step.py dam32 1 > gpurun_out/r2y_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2y_racecheck.log
timeout 900 compute-sanitizer --tool synccheck python scripts/sanitize_step.py dam32 1 > gpurun_out/r2y_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r2y_synccheck.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_step.py dam32 2 > gpurun_out/r2y_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2y_memcheck.log
for peer in 1 0; do
FLIP_PEER=$peer SLAB_CHECK_ORACLE=0 timeout 900 compute-sanitizer --tool memcheck --target-processes all python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 scripts/slab_check.py damz32 2 > gpurun_out/r2y_memcheck_slab_peer$peer.log 2>&1; echo "memcheck slab peer=$peer rc=$?"; grep -E "ERROR SUMMARY|SLAB_CHECK" gpurun_out/r2y_memcheck_slab_peer$peer.log | tail -4
done
