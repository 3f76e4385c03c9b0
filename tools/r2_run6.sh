mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q --timeout 600 2>&1 | tail -5 | cut -c1-250
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --timeout 900 -k "bf16-" 2>&1 | tail -3 | cut -c1-250
for SEG in 2 1; do
SEMSEG_B200_GRAPH_SEGMENTS=$SEG SEMSEG_B200_GRAPH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity-mode > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log | cut -c1-2500 > gpurun_out/r2_bench_2gpu_seg$SEG.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_2gpu_seg$SEG.json'))
    print('2GPU segments=$SEG', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
except Exception as e:
    print('2GPU seg=$SEG bench failed', e); print(open('gpurun_out/bench2.log').read()[-3000:])
PY
done
SEMSEG_B200_GRAPH_SEGMENTS=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/profile_ddp.py 2>&1 | grep -E "ms |rank0|total" | head -32 | cut -c1-160 > gpurun_out/r2_ddp2_step_profile_graph.txt; cat gpurun_out/r2_ddp2_step_profile_graph.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu --no-parity-mode --optimizer fused 2>&1 | tail -1 | cut -c1-300
