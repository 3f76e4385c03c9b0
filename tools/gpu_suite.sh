#!/bin/bash
# One-GPU validation suite (run through gpurun): GPU tests, bench, optional clustered-conv A/B.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu 2>&1 | tail -1 > gpurun_out/bench_plain.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_plain.json'))
print('PLAIN  ', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['clocks'])
PY
if [ "$1" = "cluster" ]; then
  SEMSEG_B200_CLUSTER=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 300 -x -k "conv or bottleneck or pspnet50_small" 2>&1 | tail -6
  SEMSEG_B200_CLUSTER=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu 2>&1 | tail -1 > gpurun_out/bench_cluster.json
  python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_cluster.json'))
    print('CLUSTER', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'])
except Exception as e:
    print('CLUSTER bench failed', e, open('gpurun_out/bench_cluster.json').read()[-400:])
PY
  SEMSEG_B200_CLUSTER=1 timeout 300 python tools/bringup.py perf 2>&1 | grep perf
  timeout 300 python tools/bringup.py perf 2>&1 | grep perf
fi
if [ "$1" = "epi" ]; then
  SEMSEG_B200_EPI_GROUPS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu 2>&1 | tail -1 > gpurun_out/bench_epi1.json
  python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_epi1.json'))
    print('EPI=1  ', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'])
except Exception as e:
    print('EPI=1 bench failed', e, open('gpurun_out/bench_epi1.json').read()[-400:])
PY
fi
python tools/profile_step.py 2>&1 | sed -n 1,24p
