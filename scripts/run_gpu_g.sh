mkdir -p gpurun_out
timeout 900 python scripts/config_table.py 100 10 2 > gpurun_out/g_configs.md 2> gpurun_out/g_configs.err; echo "configs rc=$?"; grep -v "Fluid Engine\|^---" gpurun_out/g_configs.err | tail -5; cat gpurun_out/g_configs.md
