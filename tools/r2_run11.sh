# 1-GPU: default bench line (all legs incl. the reference arms), reference arm alone, ncu refresh of the config-2 kernels
mkdir -p gpurun_out
timeout 1500 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/r2_bench_default.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_default.json'))
print('DEFAULT', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'parity', d['parity_mode'] and round(d['parity_mode']['value'],1))
print(' stock', d.get('stock_gpu_baseline'))
print(' cpu', d.get('cpu_baseline'))
print(' clocks', d['clocks'], 'launches', d['gpu_launches'])
PY
tail -3 gpurun_out/bench_default.err | cut -c1-300
timeout 900 python bench.py --impl reference --steps 4 --warmup 1 2>&1 | tail -1 | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_ncu.csv env SEMSEG_B200_GRAPH=0 python tools/profile_step.py --ncu > gpurun_out/launches_step.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/r2_launches_step_ncu.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_kernels -k regex:"conv_igemm|conv_wgrad|bn_|upsample_ce|psamask|ppm_|wgrad_reduce" python tools/ncu_kernels.py > gpurun_out/ncu_kernels.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels_raw.csv 2>/dev/null
python tools/ncu_key_metrics.py gpurun_out/prof_kernels_raw.csv > gpurun_out/r2_ncu_full_key_metrics.csv
rm -f gpurun_out/prof_kernels.ncu-rep gpurun_out/prof_kernels_raw.csv
cut -d, -f2,7,8,9,12 gpurun_out/r2_ncu_full_key_metrics.csv | cut -c1-50,150-220 | head -12
