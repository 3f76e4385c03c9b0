#!/bin/bash
# Multi-GPU validation (through `gpurun --gpus N`): multi-rank SyncBN / DDP parity against the fp32 oracle (tests/test_multigpu_gpu.py), then the weak-scaling bench at N and 2.
N=${1:-4}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
timeout 1500 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -s --timeout 900 2>&1 | grep -E "OK|FAIL|passed|failed|SyncBN|timed out" | head -40
run $N 29522 bench.py --gpus $N --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_${N}gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('P2P$N', d['config'].get('syncbn_exchange'), d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
if [ "$N" != "2" ]; then
run 2 29523 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_2gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('P2P2', d['config'].get('syncbn_exchange'), d['value'], d['ms_per_step'], d['e2e']['value'])"
fi
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stock-gpu 2>&1 | tail -1 | tee gpurun_out/bench_1gpu_samebox.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 ', d['value'], d['ms_per_step'], d['e2e']['value'])"
