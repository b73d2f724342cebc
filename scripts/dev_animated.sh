# developer batch: every test around obstacles, meshes, moving solids and friction (the code touched since the last full run)
mkdir -p gpurun_out
timeout 55 python -m pytest tests/test_moving_solids_gpu.py tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "facade or seeding or static_obstacles or moving or animated or solid_velocity or friction or stages_against_golden" > gpurun_out/r2k_obstacle_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2k_obstacle_tests.log | cut -c1-1500
