mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 600 -k "graph or sgd or psa_attend or psanet or wgrad or conv_fprop" 2>&1 | tail -6 | cut -c1-250
python tools/profile_step.py --arch psa --size 465 --batch 2 --out gpurun_out/r2_step_breakdown_psa50_465_bs2_graph.txt | grep -E "cpu enqueue|psa_|total" | cut -c1-150
for CFG in "--arch psa --size 465 --batch 2" "--arch psa --size 465 --batch 16" "--layers 101 --size 713 --classes 19 --batch 2" "--layers 101 --size 713 --classes 19 --batch 16" ""; do
  timeout 900 python bench.py $CFG --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu 2>/dev/null | tail -1 > gpurun_out/tmp.json
  python - <<PY
import json
d=json.load(open('gpurun_out/tmp.json'))
print('BENCH [$CFG] ->', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value'],1), '; x3', d['parity_mode'] and round(d['parity_mode'].get('value',0),1), '; conv TFLOP/s', d.get('step_conv_tflops'))
name='$CFG'.replace('--','').replace(' ','_') or 'config2'
open('gpurun_out/r2_bench_1gpu_'+name+'.json','w').write(json.dumps(d))
PY
done
SEMSEG_B200_N256=128 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-parity-mode 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('N256=128:', round(d['value'],1), round(d['ms_per_step'],2))"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-parity-mode --optimizer fused 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('fused sgd:', round(d['value'],1), round(d['ms_per_step'],2))"
