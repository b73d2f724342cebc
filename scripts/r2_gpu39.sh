mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --ref-budget 60 > gpurun_out/r2av_ref.json 2> gpurun_out/r2av_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/r2av_ref.json
