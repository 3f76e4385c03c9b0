mkdir -p gpurun_out
timeout 1400 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -300 > gpurun_out/r2_gpu_tests_4.log; grep -E "^(FAILED|ERROR)|passed|failed|bf16x3" gpurun_out/r2_gpu_tests_4.log | cut -c1-250
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu 2>&1 | tail -3 | cut -c1-3000 > gpurun_out/r2_bench_graph.json; cat gpurun_out/r2_bench_graph.json | cut -c1-1800
SEMSEG_B200_GRAPH=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-parity-mode 2>&1 | tail -1 | cut -c1-600
