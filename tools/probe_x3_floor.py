"""What can (hi, lo) bf16 operands achieve at best? The fp32 oracle with every conv's operands rounded to 16 mantissa
bits (bf16 hi + bf16 lo) — exact fp32 accumulation, everything else fp32 — against the plain fp32 oracle on the
PSPNet50 @ 473x473 parity case of tests/test_parity_x3_gpu.py. Separates the operand-rounding floor from what the kernels
add (accumulation order / truncation). Also a 3-limb (24-bit) variant."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tests import util
from oracle.torch_oracle import Oracle

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def limbs(t, n):
    out, r = torch.zeros_like(t), t
    for _ in range(n):
        h = r.to(torch.bfloat16).float()
        out = out + h
        r = r - h
    return out


class RoundedOracle(Oracle):
    nl = 2

    def conv(self, x, name, stride=1, padding=0, dilation=1):
        return F.conv2d(limbs(x, self.nl), limbs(self.sd[name + '.weight'], self.nl), self.sd.get(name + '.bias'),
                        stride, padding, dilation)


model = util.build_pspnet(50, 150).cuda()
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
x, y = util.synth(2, 473, 473, 150, device="cuda")
with torch.no_grad():
    lo = Oracle(sd, arch="psp", layers=50, classes=150).eval().forward(x)
    for nl in (2, 3):
        ro = RoundedOracle(sd, arch="psp", layers=50, classes=150).eval()
        ro.nl = nl
        lr = ro.forward(x)
        e = float((lr.double() - lo.double()).norm() / lo.double().norm())
        flips = int((lr.argmax(1) != lo.argmax(1)).sum())
        print("operands rounded to %d bf16 limbs, exact fp32 accumulation: rel_l2 %.3e, max abs %.3e, argmax flips %d / %d"
              % (nl, e, float((lr - lo).abs().max()), flips, lo.argmax(1).numel()), flush=True)
    # fp32 reorder noise: cudnn vs a different algorithm path (channels_last)
    lo2 = Oracle({k: v for k, v in sd.items()}, arch="psp", layers=50, classes=150).eval().forward(
        x.contiguous(memory_format=torch.channels_last))
    print("fp32 vs fp32 (channels_last kernels): rel_l2 %.3e, flips %d" %
          (float((lo2.double() - lo.double()).norm() / lo.double().norm()), int((lo2.argmax(1) != lo.argmax(1)).sum())))
