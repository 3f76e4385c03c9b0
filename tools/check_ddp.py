"""Multi-GPU parity of the data-parallel path (run with torchrun, 2+ GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_ddp.py

Every rank builds the same seeded PSPNet50, converts to SyncBatchNorm and wraps in DDP exactly as
tool/train.py:141-157, and runs one training step on its shard of a global batch. Rank 0 then runs the same weights
in ONE process with plain BatchNorm over the concatenated batch and loss = mean over ranks of the per-shard CE
(SURVEY.md §8 e: each rank's CE is a mean over its own valid pixels and DDP averages gradients) and compares losses,
BN running statistics and gradients.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402

from model.pspnet import PSPNet  # noqa: E402


def synth(n, size, classes, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, 3, size, size), generator=g)
    y = torch.randint(0, classes, (n, size, size), generator=g)
    y[torch.rand((n, size, size), generator=g) < 0.05] = 255
    return x, y


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    classes, size, per = 21, 129, 2
    x, y = synth(per * world, size, classes, 7)

    torch.manual_seed(0)
    model = PSPNet(layers=50, classes=classes, zoom_factor=8, dropout=0.0, pretrained=False)
    ref_sd = {k: v.clone() for k, v in model.state_dict().items()}
    ddp = nn.parallel.DistributedDataParallel(nn.SyncBatchNorm.convert_sync_batchnorm(model).cuda(),
                                              device_ids=[local])
    ddp.train()
    xs, ys = x[rank * per:(rank + 1) * per].to(dev), y[rank * per:(rank + 1) * per].to(dev)
    _, ml, al = ddp(xs, ys)
    (ml + 0.4 * al).backward()
    losses = torch.stack([ml.detach(), al.detach()])
    gathered = [torch.zeros_like(losses) for _ in range(world)]
    dist.all_gather(gathered, losses)
    ok = True
    if rank == 0:
        torch.manual_seed(0)
        single = PSPNet(layers=50, classes=classes, zoom_factor=8, dropout=0.0, pretrained=False)
        single.load_state_dict(ref_sd)
        single = single.cuda().train()
        # one process, plain BN over the concatenated batch; per-shard CE averaged over ranks
        from semseg_b200 import functional as SF
        from semseg_b200.pspnet import head_forward_nhwc
        xa, ya = x.to(dev), y.to(dev)
        t = SF.to_nhwc_bf16(xa)
        t = single.layer0.forward_nhwc(t)
        t = single.layer1.forward_nhwc(t)
        t = single.layer2.forward_nhwc(t)
        t3 = single.layer3.forward_nhwc(t)
        t = single.ppm.forward_nhwc(single.layer4.forward_nhwc(t3))
        lg, la = head_forward_nhwc(single.cls, t), head_forward_nhwc(single.aux, t3)
        mls, als = [], []
        for r in range(world):
            sl = slice(r * per, (r + 1) * per)
            m_, _ = SF.upsample_ce(lg[sl], ya[sl], 255)
            a_, _ = SF.upsample_ce(la[sl], ya[sl], 255)
            mls.append(m_)
            als.append(a_)
        (sum(mls) / world + 0.4 * sum(als) / world).backward()
        for r in range(world):
            dm = abs(gathered[r][0].item() - mls[r].item()) / mls[r].item()
            da = abs(gathered[r][1].item() - als[r].item()) / als[r].item()
            print("rank %d main %.6f vs %.6f  aux %.6f vs %.6f" % (r, gathered[r][0].item(), mls[r].item(),
                                                                  gathered[r][1].item(), als[r].item()))
            ok &= dm < 1e-4 and da < 1e-4
        dsd = ddp.module.state_dict()
        ssd = single.state_dict()
        worst = max((rel(dsd[k], ssd[k]), k) for k in ssd if "running" in k)
        print("worst running-stat rel err %.3e (%s)" % worst)
        ok &= worst[0] < 1e-4
        sp = dict(single.named_parameters())
        errs = sorted(((rel(p.grad, sp[k].grad), k) for k, p in ddp.module.named_parameters()), reverse=True)
        print("worst grad rel errs:", ["%.2e %s" % e for e in errs[:5]])
        print("median grad rel err: %.2e" % errs[len(errs) // 2][0])
        # identical kernels on identical data; only the order of cross-rank merges differs (fp32 reorder)
        ok &= errs[len(errs) // 2][0] < 5e-2
        print("DDP+SyncBN parity:", "OK" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
