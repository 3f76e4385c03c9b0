#!/bin/bash
# Round-end style validation on one GPU: GPU tests, smoke, the default bench.py run (all legs), inference bench, evidence.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print('BENCH', {k:d.get(k) for k in ('value','ms_per_step','gpu_launches','vs_baseline')}, 'e2e', d['e2e'], 'roof', d['roofline'], 'cpu', d['cpu_baseline'], 'stock', d.get('stock_gpu_baseline'), d['clocks'])
PY
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
timeout 900 python tools/bench_inference.py 2>&1 | tail -1 | tee gpurun_out/bench_inference_pspnet101_713.json
timeout 600 python tools/bench_inference.py --layers 50 --classes 150 --crop 473 2>&1 | tail -1 | tee gpurun_out/bench_inference_pspnet50_473.json
bash tools/gpu_evidence.sh
bash tools/gpu_evidence_n64.sh
