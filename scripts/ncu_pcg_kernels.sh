mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_mg_coarse|k_pcg_update_presweep|k_mg0_restrict_list|k_mg0_sweep|k_pcg_dir_spmv" -c 6 --launch-skip 60 -o gpurun_out/r2bb_pcg -f python bench.py --steps 1 --warmup 1 --no-parity-check --cpu-budget 0 --exact-steps 0 > gpurun_out/r2bb.log 2>&1; echo "rc=$?"
python scripts/ncu_brief.py gpurun_out/r2bb_pcg.ncu-rep > gpurun_out/r2bb_pcg_digest.txt; wc -l gpurun_out/r2bb_pcg_digest.txt
