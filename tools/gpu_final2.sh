#!/bin/bash
# Evidence pass: default bench JSON, inference bench (config-5 shape), ncu launch list + full capture (CSV exports only).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 -k "sliding or pack or ppm" 2>&1 | tail -3
python bench.py 2>gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.json
python -c "import json; d=json.load(open('gpurun_out/bench_default.json')); print('BENCH', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
timeout 900 python tools/bench_inference.py 2>&1 | tail -1 | tee gpurun_out/bench_inference_pspnet101_713.json
bash tools/gpu_evidence.sh
bash tools/gpu_evidence_n64.sh
du -sh gpurun_out
