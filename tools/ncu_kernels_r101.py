"""Launch the ResNet-101 dilated-stage kernels (BASELINE config 4: PSPNet101 @ 713x713 -> 90x90 maps) once each, for
`ncu --set full` captures of the tensor-pipe utilisation north_star targets (>= 70 % on the dilated stage):

    ncu --set full --clock-control none --import-source on -o gpurun_out/prof_r101 python tools/ncu_kernels_r101.py [N]

N = images per GPU (2 = the reference's global batch 16 on 8 GPUs, tool/train.py:154; 16 = the weak-scaling variant).
Layer list = SURVEY.md Appendix D, PSPNet101 @713: layer3 3x3 d2 256->256 (x23), its 1x1 neighbours 256->1024 /
1024->256, layer4 3x3 d4 512->512, 512->2048, 2048->512, the cls 3x3 4096->512; fprop, dgrad and wgrad of each.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from semseg_b200 import ops  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    split = len(sys.argv) > 2 and sys.argv[2] == "x3"
    hw = 90
    g = torch.Generator(device="cuda").manual_seed(0)

    def act(c):
        t = torch.randn((n, hw, hw, c), device="cuda", generator=g)
        return ops.f32_to_act(t, True) if split else t.to(torch.bfloat16)

    layers = [(256, 256, 3, 2), (256, 1024, 1, 1), (1024, 256, 1, 1), (512, 512, 3, 4), (512, 2048, 1, 1),
              (2048, 512, 1, 1), (4096, 512, 3, 1)]
    for cin, cout, k, dil in layers:
        x, dy = act(cin), act(cout)
        w = torch.randn((cout, cin, k, k), device="cuda", generator=g) * 0.02
        pw = ops.pack_weights(w, split=split)
        torch.cuda.synchronize()
        ops.conv_fprop(x, pw.wf, cout, ops.conv_taps(k, dil), stats=True)
        ops.conv_fprop(dy, pw.wd, cin, ops.conv_taps(k, dil, transpose=True))
        ops.conv_wgrad(x, dy, cin, cout, ops.conv_taps(k, dil))
        torch.cuda.synchronize()
        del x, dy, w, pw


if __name__ == "__main__":
    main()
