"""Per-kernel breakdown of one DDP + SyncBN training step on rank 0 (run under torchrun with >= 2 GPUs)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402

import bench  # noqa: E402
from model.pspnet import PSPNet  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = PSPNet(layers=50, classes=150, zoom_factor=8, pretrained=False)
    opt = bench.build_optimizer(model, "psp")
    model = nn.parallel.DistributedDataParallel(nn.SyncBatchNorm.convert_sync_batchnorm(model).cuda(),
                                                device_ids=[local])
    model.train()
    x, y = bench.synth_batch(16, 473, 150, 100 + rank)
    x, y = x.to(dev), y.to(dev)

    def step():
        _, ml, al = model(x, y)
        loss = ml + 0.4 * al
        opt.zero_grad()
        loss.backward()
        opt.step()

    for _ in range(9):          # eager warm-up + capture of the step graphs (unless SEMSEG_B200_GRAPH=0)
        step()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    t_cpu = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    t_gpu = e0.elapsed_time(e1) / 5
    if rank == 0:
        print("rank0: cpu enqueue %.1f ms/step, gpu %.1f ms/step" % (t_cpu * 1e3, t_gpu), flush=True)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    if rank == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
        agg = {}
        for e in ev:
            a = agg.setdefault(e.name, [0.0, 0])
            a[0] += e.device_time
            a[1] += 1
        tot = sum(v[0] for v in agg.values())
        print("total kernel time %.2f ms over %d launches" % (tot / 1e3, sum(v[1] for v in agg.values())))
        for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
            print("%8.3f ms %5.1f%% x%-4d %s" % (t / 1e3, 100 * t / tot, c, name[:100]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
