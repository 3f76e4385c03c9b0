# 8-GPU evidence: BASELINE configs 2 / 3 / 4 (reference batch split 2 images/GPU and the 16/GPU weak variants), parity
mkdir -p gpurun_out
run() {  # name, bench args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 --no-parity-mode $2 > gpurun_out/bench8.log 2>&1
  tail -1 gpurun_out/bench8.log > gpurun_out/r2_bench_8gpu_$1.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2_bench_8gpu_$1.json'))
    print('8GPU $1:', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms/step; e2e', round(d['e2e']['value'],1), d['config'].get('syncbn_exchange'), d['clocks'])
except Exception as e:
    print('8GPU $1 failed', e); print(open('gpurun_out/bench8.log').read()[-2500:])
PY
}
run config2_psp50_473_bs16 ""
run config3_psa50_465_bs2 "--arch psa --size 465 --batch 2"
run config4_psp101_713_bs2 "--layers 101 --size 713 --classes 19 --batch 2"
run config3w_psa50_465_bs16 "--arch psa --size 465 --batch 16"
run config4w_psp101_713_bs16 "--layers 101 --size 713 --classes 19 --batch 16"
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s --timeout 900 -k eight 2>&1 | grep -E "OK|FAIL|passed|failed|SyncBN|rank 0 main" | cut -c1-200 | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/profile_ddp.py 2>&1 | grep -E "ms |rank0|total" | head -30 | cut -c1-160 > gpurun_out/r2_ddp8_step_profile_graph.txt; head -30 gpurun_out/r2_ddp8_step_profile_graph.txt
