mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q --timeout 600 2>&1 | tail -120 | cut -c1-220 > gpurun_out/graph_tests.log; grep -E "^E  |passed|failed|Error" gpurun_out/graph_tests.log | head -40
timeout 600 python tools/bench_psamask.py 2>&1 | tail -14 | cut -c1-260 > gpurun_out/r2_psamask_vs_stock.txt; head -12 gpurun_out/r2_psamask_vs_stock.txt
bash tools/r2_run3.sh
