# round-end style validation on one B200: tests, bench, ncu launch list, ncu full captures of the main kernels
mkdir -p gpurun_out
T=${1:-r}
timeout 1200 python -m pytest tests -m gpu -x -q --tb=short > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv python scripts/profile_step.py sphere256 3 1 > gpurun_out/${T}_prof.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"k_sdf_p2g|k_sdf_far|k_g2p_fast|k_advance_fast|k_gather|k_classify|k_cell_finalize|k_scatter_idx|k_build_src|k_speed_hist|k_build_system|k_apply_pressure|k_seg_flag|k_ext_init" -c 16 \
  -f -o gpurun_out/${T}_particles python scripts/profile_step.py sphere256 3 1 > gpurun_out/${T}_particles.log 2>&1; echo "ncu A rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"k_pcg_spmv|k_pcg_update|k_pcg_direction|k_mg0_sweep|k_mg0_restrict_list|k_mg_coarse|k_ext_claim|k_ext_fill" --launch-skip 16 -c 10 \
  -f -o gpurun_out/${T}_pcg python scripts/profile_step.py sphere256 3 1 > gpurun_out/${T}_pcg.log 2>&1; echo "ncu B rc=$?"
python - <<P
import json
d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'], 'scaling_ref', d.get('scaling_ref'))
print('roofline', d['roofline']); print('cpu', d['cpu_baseline'])
for k,v in d['kernels'].items(): print(k, round(v['avg_ms'],4), v['launches'], round(v.get('frac',0),3))
print(open('gpurun_out/${T}_bench_ref.json').read()[:600])
P
