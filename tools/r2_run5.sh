mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_graph_gpu.py -m gpu -q --timeout 600 -k "resize or graphed or psanet or psa_attend" 2>&1 | tail -15 | cut -c1-250
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s --timeout 900 2>&1 | grep -E "OK|FAIL|passed|failed|rank [01] main|grad |worst|ratio|SyncBN" | cut -c1-220 | tail -70 > gpurun_out/r2_multigpu_tests.log; cat gpurun_out/r2_multigpu_tests.log
for G in 1 0; do
SEMSEG_B200_GRAPH=$G SEMSEG_B200_GRAPH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity-mode > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log | cut -c1-2500 > gpurun_out/r2_bench_2gpu_graph$G.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_2gpu_graph$G.json'))
    print('2GPU graph=$G', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config'].get('execution'), d['config'].get('syncbn_exchange'))
except Exception as e:
    print('2GPU graph=$G bench failed', e); print(open('gpurun_out/bench2.log').read()[-3000:])
PY
done
timeout 1500 python tools/run_reference_trainer.py --gpus 2 --iters 20 --out gpurun_out/trainer 2>&1 | tail -12 | cut -c1-300
