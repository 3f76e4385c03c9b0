# 2-GPU check of the push-model SyncBN exchange: parity tests + DDP bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s --timeout 900 2>&1 | grep -E "OK|FAIL|passed|failed|SyncBN" | cut -c1-200 | tail -30
SEMSEG_B200_GRAPH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity-mode > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log > gpurun_out/r2_bench_2gpu.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_2gpu.json'))
    print('2GPU', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
except Exception as e:
    print('2GPU bench failed', e); print(open('gpurun_out/bench2.log').read()[-3000:])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/profile_ddp.py 2>&1 | grep -E "ms |rank0|total" | head -24 | cut -c1-160 > gpurun_out/r2_ddp2_step_profile_graph.txt; grep -E "rank0|p2p|Memcpy|nccl|AUnary" gpurun_out/r2_ddp2_step_profile_graph.txt
