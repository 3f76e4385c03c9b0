"""Worker of tests/test_multigpu_gpu.py (launched with torch.distributed.run on >= 2 GPUs).

Multi-rank SyncBatchNorm / DDP parity against the fp32 ORACLE (oracle/torch_oracle.py: plain torch fp32, TF32 off) run
by rank 0 in one process over the CONCATENATED batch — not against this repository's own single-process kernels:

  * blocks (Bottleneck with downsample d=2, Bottleneck d=4, PPM) under nn.SyncBatchNorm, one shard per rank:
      outputs, input gradients, summed parameter gradients and BN running statistics vs the oracle on the whole batch.
  * PSPNet50 wrapped exactly as tool/train.py:141-157 (convert_sync_batchnorm + DistributedDataParallel): per-rank losses
      vs the oracle's per-shard cross-entropy (SURVEY.md §8 e: each rank's CE is a mean over its own valid pixels),
      running statistics, and DDP-averaged head gradients vs d/dθ of mean_r(loss_r).

Tolerances depend on the operand mode (argv[1]): bf16x3 -> 1e-4 forward, 1e-5 running statistics, gradients to the ReLU
mask-flip floor (1e-2 against the plain oracle; the mask-matched 1e-3 check lives in tests/test_parity_x3_gpu.py);
bf16 -> the single-pass bf16 floors used by tests/test_parity_gpu.py. Prints one line per check and exits non-zero on any failure.
"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


def gather_cat(t):
    t = t.contiguous()
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.cat(out)


def block_checks(rank, world, dev, mode):
    from semseg_b200 import functional as SF, ops
    from semseg_b200.resnet import Bottleneck
    from semseg_b200.pspnet import PPM
    from oracle.torch_oracle import Oracle
    x3 = mode == "bf16x3"
    # gradients against the PLAIN fp32 oracle: the ReLU mask-flip floor sqrt(0.8 * forward error) applies
    # (tests/test_parity_x3_gpu.py::test_bottleneck_block_x3_vs_oracle), ~3e-3 at a forward error of 1e-5
    tol_y, tol_g, tol_run = (1e-4, 1e-2, 1e-5) if x3 else (8e-3, 0.15, 1e-2)
    ok, per = True, 2
    torch.manual_seed(1)
    ds = nn.Sequential(nn.Conv2d(256, 512, 1, bias=False), nn.BatchNorm2d(512))
    cases = [("bottleneck+downsample d2", Bottleneck(256, 128, 1, ds), (256, 30, 30), 2),
             ("bottleneck d4", Bottleneck(512, 128), (512, 30, 30), 4),
             ("ppm", PPM(256, 64, (1, 2, 3, 6)), (256, 12, 12), 0)]
    for name, mod, (c, h, w), dil in cases:
        if dil:
            mod.conv2.dilation, mod.conv2.padding = (dil, dil), (dil, dil)
        for m in mod.modules():
            if isinstance(m, nn.BatchNorm2d):
                nn.init.uniform_(m.weight, 0.5, 1.5)
                nn.init.normal_(m.bias, 0, 0.2)
        g = torch.Generator().manual_seed(11)
        x = torch.randn((per * world, c, h, w), generator=g)
        sync = nn.SyncBatchNorm.convert_sync_batchnorm(copy.deepcopy(mod)).to(dev).train()
        for p_s in sync.parameters():
            dist.broadcast(p_s.data, 0)
        ref_sd = {k: v.detach().clone() for k, v in sync.state_dict().items()}     # weights before the step
        xs = SF.to_nhwc_bf16(x[rank * per:(rank + 1) * per].to(dev)).requires_grad_(True)
        ys = sync.forward_nhwc(xs)
        out_c = ys.shape[-1]
        gfull = torch.randn((per * world, h, w, out_c), generator=g)
        gy = ops.f32_to_act(gfull[rank * per:(rank + 1) * per].to(dev).contiguous(), ops.is_split(ys)) \
            if ops.is_split(ys) else gfull[rank * per:(rank + 1) * per].to(dev).to(torch.bfloat16)
        ys.backward(gy)
        for p_s in sync.parameters():
            dist.all_reduce(p_s.grad)                       # sum over ranks = gradient of the whole-batch objective
        y_all = gather_cat(ops.act_to_f32(ys.detach()))
        dx_all = gather_cat(ops.act_to_f32(xs.grad))
        x_all = gather_cat(ops.act_to_f32(xs.detach()))      # exactly what the kernels saw (bf16 / hi+lo rounding of x)
        g_all = gather_cat(ops.act_to_f32(gy))
        if rank == 0:
            # fp32 oracle, ONE process, plain BatchNorm over the concatenated batch
            sd = {("layer1.0." if dil else "ppm.") + k: v.clone() for k, v in ref_sd.items()}
            for k, v in sd.items():
                if v.dtype.is_floating_point and "running" not in k:
                    v.requires_grad_(True)
            orc = Oracle(sd, bins=(1, 2, 3, 6))
            xo = x_all.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
            yo = orc.bottleneck(xo, "layer1.0", 1, dil, name.startswith("bottleneck+")) if dil else orc.ppm(xo)
            yo.backward(g_all.permute(0, 3, 1, 2))
            e_y = rel(y_all.permute(0, 3, 1, 2), yo)
            e_dx = rel(dx_all.permute(0, 3, 1, 2), xo.grad)
            pre = "layer1.0." if dil else "ppm."
            sp = dict(sync.named_parameters())
            skip = (lambda k: "features.0" in k) if not dil else (lambda k: False)   # PPM bin 1: ~0/0 gradients
            e_p = max(rel(p.grad, sd[pre + k].grad) for k, p in sp.items() if not skip(k))
            c_p = min(cos(p.grad, sd[pre + k].grad) for k, p in sp.items() if not skip(k))
            bs = dict(sync.named_buffers())
            e_b = max(rel(bs[k], sd[pre + k]) for k in bs if "running" in k)
            good = e_y < tol_y and e_dx < tol_g and e_p < tol_g and e_b < tol_run and (x3 or c_p > 0.99)
            print("%-26s [%s] y %.2e  dx %.2e  dparam %.2e (cos %.4f)  running %.2e  %s" %
                  (name, mode, e_y, e_dx, e_p, c_p, e_b, "OK" if good else "FAIL"), flush=True)
            ok &= good
    return ok


def network_check(rank, world, local, dev, mode):
    from model.pspnet import PSPNet
    from oracle.torch_oracle import Oracle
    x3 = mode == "bf16x3"
    classes, size, per = 21, 129, 2
    g = torch.Generator().manual_seed(7)
    x = torch.randn((per * world, 3, size, size), generator=g)
    y = torch.randint(0, classes, (per * world, size, size), generator=g)
    y[torch.rand((per * world, size, size), generator=g) < 0.05] = 255
    torch.manual_seed(0)
    model = PSPNet(layers=50, classes=classes, zoom_factor=8, dropout=0.0, pretrained=False)
    ref_sd = {k: v.clone() for k, v in model.state_dict().items()}
    ddp = nn.parallel.DistributedDataParallel(nn.SyncBatchNorm.convert_sync_batchnorm(model).cuda(),
                                              device_ids=[local])                 # tool/train.py:141-157
    ddp.train()
    xs, ys = x[rank * per:(rank + 1) * per].to(dev), y[rank * per:(rank + 1) * per].to(dev)
    _, ml, al = ddp(xs, ys)
    (ml + 0.4 * al).backward()
    losses = gather_cat(torch.stack([ml.detach(), al.detach()]).view(1, 2))
    ok = True
    if rank == 0:
        params = {k for k, _ in model.named_parameters()}
        sd = {k: v.clone().to(dev) for k, v in ref_sd.items()}
        for k, v in sd.items():
            if k in params:
                v.requires_grad_(True)
        orc = Oracle(sd, arch="psp", layers=50, classes=classes).train()
        xa, ya = x.to(dev), y.to(dev)
        main, aux = orc.logits_lowres(xa)                    # plain BN over the concatenated batch
        main = F.interpolate(main, size=(size, size), mode="bilinear", align_corners=True)
        aux = F.interpolate(aux, size=(size, size), mode="bilinear", align_corners=True)
        mls, als = [], []
        for r in range(world):
            sl = slice(r * per, (r + 1) * per)
            mls.append(F.cross_entropy(main[sl], ya[sl], ignore_index=255))
            als.append(F.cross_entropy(aux[sl], ya[sl], ignore_index=255))
        (sum(mls) / world + 0.4 * sum(als) / world).backward()       # DDP averages the per-rank gradients
        tol_l = 1e-4 if x3 else 2e-3
        for r in range(world):
            dm = abs(losses[r][0].item() - mls[r].item()) / mls[r].item()
            da = abs(losses[r][1].item() - als[r].item()) / als[r].item()
            print("rank %d main %.6f vs oracle %.6f  aux %.6f vs %.6f" % (r, losses[r][0].item(), mls[r].item(),
                                                                         losses[r][1].item(), als[r].item()))
            ok &= dm < tol_l and da < tol_l
        dsd = ddp.module.state_dict()
        worst = max((rel(dsd[k], sd[k]), k) for k in sd if "running" in k)
        print("worst running-stat rel err vs oracle %.3e (%s)" % worst)
        ok &= worst[0] < (1e-3 if x3 else 1e-1)
        # the classifier layers sit downstream of everything: their DDP-averaged gradients vs the oracle
        dp = dict(ddp.module.named_parameters())
        for k in ("cls.4.weight", "cls.4.bias", "aux.4.weight", "aux.4.bias"):
            e = rel(dp[k].grad, sd[k].grad)
            print("grad %-14s rel err %.3e" % (k, e))
            ok &= e < (2e-2 if x3 else 0.3)
        norms = sorted((float(dp[k].grad.double().norm()) / (float(sd[k].grad.double().norm()) + 1e-30), k)
                       for k in dp if float(sd[k].grad.double().norm()) > 1e-6 and "ppm.features.0" not in k)
        print("gradient norm ratio vs oracle: min %.3f (%s)  max %.3f (%s)" % (norms[0] + norms[-1]))
        ok &= 0.5 < norms[0][0] and norms[-1][0] < 2.0
    return ok


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dist.init_process_group("nccl", device_id=dev)
    from semseg_b200 import precision, p2p
    precision.set_mode(mode)
    ok = block_checks(rank, world, dev, mode)
    ok = network_check(rank, world, local, dev, mode) and ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SyncBN exchange: %s" % p2p.exchange_kind())
        print("multi-rank parity [%s, world %d]: %s" % (mode, world, "OK" if flag.item() else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
