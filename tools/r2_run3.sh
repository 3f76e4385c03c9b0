# 1-GPU evidence for BASELINE configs 3 / 4 and the ResNet-101 dilated stage (N2 / N3 of VERDICT r1).
mkdir -p gpurun_out
# ncu --set full of the R101 dilated-stage kernels at 90x90: 2 images/GPU (reference batch split) and 16 images/GPU
for N in 2 16; do
  timeout 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof_r101_n$N \
    -k regex:"conv_igemm|conv_wgrad" python tools/ncu_kernels_r101.py $N > gpurun_out/ncu_r101_n$N.log 2>&1
  echo "ncu r101 n=$N rc=$?"
  ncu -i gpurun_out/prof_r101_n$N.ncu-rep --page raw --csv > gpurun_out/prof_r101_n${N}_raw.csv 2>/dev/null
  python tools/ncu_key_metrics.py gpurun_out/prof_r101_n${N}_raw.csv > gpurun_out/r2_ncu_r101_n${N}_key_metrics.csv
  rm -f gpurun_out/prof_r101_n$N.ncu-rep gpurun_out/prof_r101_n${N}_raw.csv
done
cut -d, -f2,7,12 gpurun_out/r2_ncu_r101_n2_key_metrics.csv | cut -c1-160 | head -30
# step breakdowns (kineto): config 4 / config 3 per-GPU shards, eager (host enqueue visible) and graph replay
SEMSEG_B200_GRAPH=0 python tools/profile_step.py --layers 101 --size 713 --classes 19 --batch 2 --out gpurun_out/r2_step_breakdown_psp101_713_bs2_eager.txt | head -3
python tools/profile_step.py --layers 101 --size 713 --classes 19 --batch 2 --out gpurun_out/r2_step_breakdown_psp101_713_bs2_graph.txt | head -30
SEMSEG_B200_GRAPH=0 python tools/profile_step.py --arch psa --size 465 --batch 2 --out gpurun_out/r2_step_breakdown_psa50_465_bs2_eager.txt | head -3
python tools/profile_step.py --arch psa --size 465 --batch 2 --out gpurun_out/r2_step_breakdown_psa50_465_bs2_graph.txt | head -40
python tools/profile_step.py --arch psa --size 465 --batch 16 --out gpurun_out/r2_step_breakdown_psa50_465_bs16_graph.txt | head -3
python tools/profile_step.py --out gpurun_out/r2_step_breakdown_psp50_473_bs16_graph.txt | head -30
# single-GPU bench lines of configs 3 / 4 (per-GPU shard 2 images, and the 16-image weak variant)
for CFG in "--arch psa --size 465 --batch 2" "--arch psa --size 465 --batch 16" "--layers 101 --size 713 --classes 19 --batch 2" "--layers 101 --size 713 --classes 19 --batch 16"; do
  timeout 900 python bench.py $CFG --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu 2>&1 | tail -1 | cut -c1-1500 > gpurun_out/tmp.json
  python - <<PY
import json
d=json.load(open('gpurun_out/tmp.json'))
print('BENCH $CFG ->', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value'],1), '; x3', d['parity_mode'] and round(d['parity_mode'].get('value',0),1), '; conv TFLOP/s', d.get('step_conv_tflops'))
open('gpurun_out/r2_bench_1gpu_'+'$CFG'.replace('--','').replace(' ','_')+'.json','w').write(json.dumps(d))
PY
done
