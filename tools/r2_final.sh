# Final single-GPU validation of the tree state (GPU budget nearly spent: keep it short)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_graph_gpu.py -m gpu -q --timeout 90 2>&1 | tail -3 | cut -c1-200
timeout 200 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/r2_bench_default.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_default.json'))
    print('DEFAULT', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'parity', d['parity_mode'] and round(d['parity_mode'].get('value',0),1))
    print(' stock', d.get('stock_gpu_baseline'))
    print(' cpu', d.get('cpu_baseline'))
    print(' clocks', d['clocks'], 'launches', d['gpu_launches'])
except Exception as e:
    print('default bench failed', e); print(open('gpurun_out/bench_default.err').read()[-2000:])
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
