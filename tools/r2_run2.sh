# 2-GPU validation: multi-rank parity tests, DDP bench with and without step graphs, the unmodified reference trainer,
# psa_mask against the stock kernel. Run with: gpurun --gpus 2 -- bash tools/r2_run2.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q --timeout 600 2>&1 | tail -15
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s --timeout 900 2>&1 | tail -80 > gpurun_out/r2_multigpu_tests.log; grep -E "OK|FAIL|passed|failed|rank|grad |worst|ratio" gpurun_out/r2_multigpu_tests.log | cut -c1-200 | tail -60
for G in 1 0; do
SEMSEG_B200_GRAPH=$G timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity-mode 2>&1 | tail -1 | cut -c1-2500 > gpurun_out/r2_bench_2gpu_graph$G.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_2gpu_graph$G.json'))
print('2GPU graph=$G', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config'].get('execution'), d['config'].get('syncbn_exchange'))
PY
done
timeout 1500 python tools/run_reference_trainer.py --gpus 2 --iters 20 --out gpurun_out/trainer 2>&1 | tail -12
timeout 600 python tools/bench_psamask.py 2>&1 | tail -6 | cut -c1-400
