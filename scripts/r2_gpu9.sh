mkdir -p gpurun_out
for v in 0 1; do
FLIP_P2G_VARIANT=$v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:"k_p2g|k_sdf|k_occ" --csv --log-file gpurun_out/r2i_l$v.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/r2i_prof$v.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/r2i_l$v.csv gpurun_out/r2i_l$v.md; echo "variant $v"; grep -E "p2g_scatter" gpurun_out/r2i_l$v.md
done
FLIP_P2G_VARIANT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or lockstep_isolated or chained or restatement or dambreak128_developed" 2>&1 | tail -3
