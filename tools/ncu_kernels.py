"""Launch each hot kernel once at its config-2 (PSPNet50, bs16, 473x473) shape, for `ncu --set full` captures.

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_kernels python tools/ncu_kernels.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from semseg_b200 import ops  # noqa: E402


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    n, hw = 16, 60

    def act(c, h=hw):
        return torch.randn((n, h, h, c), device="cuda", generator=g).to(torch.bfloat16)

    # cls head 3x3 4096->512 fprop (the dominant kernel) and layer4 conv2 (3x3 d4 512->512) fprop/dgrad/wgrad
    x4096 = act(4096)
    w_cls = torch.randn((512, 4096, 3, 3), device="cuda", generator=g) * 0.01
    pw = ops.pack_weights(w_cls)
    torch.cuda.synchronize()
    ops.conv_fprop(x4096, pw.wf, 512, ops.conv_taps(3, 1), stats=True)
    x512, dy512 = act(512), act(512)
    w4 = torch.randn((512, 512, 3, 3), device="cuda", generator=g) * 0.02
    pw4 = ops.pack_weights(w4)
    ops.conv_fprop(x512, pw4.wf, 512, ops.conv_taps(3, 4), stats=True)
    ops.conv_fprop(dy512, pw4.wd, 512, ops.conv_taps(3, 4, transpose=True))
    ops.conv_wgrad(x512, dy512, 512, 512, ops.conv_taps(3, 4))
    # 1x1 512->2048 (epilogue heavy)
    w1 = torch.randn((2048, 512, 1, 1), device="cuda", generator=g) * 0.05
    pw1 = ops.pack_weights(w1)
    raw, sp = ops.conv_fprop(x512, pw1.wf, 2048, ops.conv_taps(1, 1), stats=True)
    # 1x1 2048->512 (K = 2048) and the layer1 3x3 64->64 at 119x119 (multi-tap wgrad units)
    x2048 = act(2048)
    w2 = torch.randn((512, 2048, 1, 1), device="cuda", generator=g) * 0.02
    ops.conv_fprop(x2048, ops.pack_weights(w2).wf, 512, ops.conv_taps(1, 1), stats=True)
    x64, dy64 = act(64, 119), act(64, 119)
    w64 = torch.randn((64, 64, 3, 3), device="cuda", generator=g) * 0.05
    pw64 = ops.pack_weights(w64)
    ops.conv_fprop(x64, pw64.wf, 64, ops.conv_taps(3, 1), stats=True)
    ops.conv_wgrad(x64, dy64, 64, 64, ops.conv_taps(3, 1))
    # BN kernels on [16,60,60,2048]
    gamma = torch.ones(2048, device="cuda")
    beta = torch.zeros(2048, device="cuda")
    mi, ss = ops.bn_finalize_partials(sp, gamma, beta, 1e-5, 0.1, None, None)
    res = act(2048)
    y = ops.bn_apply(raw, ss, residual=res, relu=True)
    dy = act(2048)
    sums = ops.bn_bwd_reduce(dy, y, raw, mi, True)
    ops.bn_bwd_apply(dy, y, raw, mi, gamma, sums, float(n * hw * hw), True, want_dres=True)
    # no-residual form: ReLU mask recomputed from the raw conv output, y is never read
    sums2 = ops.bn_bwd_reduce(dy, None, raw, mi, True, scale_shift=ss)
    ops.bn_bwd_apply(dy, None, raw, mi, gamma, sums2, float(n * hw * hw), True, scale_shift=ss)
    # fused tail
    logits = torch.randn((n, hw, hw, 150), device="cuda", generator=g)
    target = torch.randint(0, 150, (n, 473, 473), device="cuda", generator=g)
    info, am, lse = ops.upsample_ce_fwd(logits, target, 255)
    ops.upsample_ce_bwd(logits, target, 255, lse, info, torch.ones((), device="cuda"))
    # PPM
    pooled = ops.ppm_pool(res, (1, 2, 3, 6))
    feats = [torch.randn((n, b, b, 512), device="cuda", generator=g).to(torch.bfloat16) for b in (1, 2, 3, 6)]
    out = ops.ppm_upsample_concat(res, feats, (1, 2, 3, 6))
    ops.ppm_upsample_bwd(out, 2048, (1, 2, 3, 6), 512)
    # psa_mask at config-3 per-GPU size and at bs16
    xm = torch.randn((16, 59 * 59, 30, 30), device="cuda", generator=g)
    col = ops.psamask_fwd(xm, 0, 59, 59)
    ops.psamask_fwd(xm, 1, 59, 59)
    ops.psamask_bwd(col, 0, 59, 59)
    ops.psamask_bwd(col, 1, 59, 59)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
