mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 600 -k "psa_attend or psamask" 2>&1 | tail -30 | cut -c1-250
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q --timeout 600 2>&1 | tail -15 | cut -c1-250
timeout 600 python tools/bench_psamask.py 2>&1 | tail -6 | cut -c1-300
