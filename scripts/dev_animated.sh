# developer batch: the tests around the obstacle / mesh code (facade examples, seeding of meshes, static and moving solids)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_moving_solids_gpu.py -m gpu -x -q --tb=short -k "facade or seeding or static_obstacles or moving or animated or solid_velocity" > gpurun_out/r2i_obstacle_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2i_obstacle_tests.log | cut -c1-1500
build/obstacles_and_sources 40 40 | tail -3
