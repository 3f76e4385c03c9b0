"""Shared helpers for the tests (synthetic inputs identical to tests/golden/_ref_worker.py)."""
import numpy as np
import torch


def synth(n, h, w, classes, seed=123, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, 3, h, w), generator=g)
    y = torch.randint(0, classes, (n, h, w), generator=g)
    ign = torch.rand((n, h, w), generator=g) < 0.05
    y[ign] = 255
    return x.to(device), y.to(device)


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a.detach()).double().flatten().cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b.detach()).double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def build_pspnet(layers=50, classes=150, seed=0, **kw):
    from semseg_b200.pspnet import PSPNet
    torch.manual_seed(seed)
    return PSPNet(layers=layers, classes=classes, zoom_factor=8, dropout=0.0, pretrained=False, **kw)


def build_psanet(layers=50, classes=150, seed=0, mask=9, **kw):
    from semseg_b200.psanet import PSANet
    torch.manual_seed(seed)
    return PSANet(layers=layers, classes=classes, zoom_factor=8, dropout=0.0, psa_type=2, compact=False,
                  shrink_factor=2, mask_h=mask, mask_w=mask, pretrained=False, **kw)


def oracle_from(model, arch, **kw):
    """fp32 oracle sharing (clones of) the model's parameters and buffers."""
    from oracle.torch_oracle import Oracle
    params = {k for k, _ in model.named_parameters()}
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    for k, v in sd.items():
        if k in params:
            v.requires_grad_(True)
    return Oracle(sd, arch=arch, **kw), sd


class TinySegNet(torch.nn.Module):
    """Deterministic stand-in network of the sliding-window goldens (identical to tests/golden/_ref_worker.py)."""

    def __init__(self, classes=5, stride=1, seed=5):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.weight = torch.nn.Parameter(torch.randn((classes, 3, 3, 3), generator=g) * 0.6)
        self.bias = torch.nn.Parameter(torch.randn((classes,), generator=g) * 0.1)
        self.stride = stride

    def forward(self, x):
        if self.stride > 1:
            x = torch.nn.functional.avg_pool2d(x, self.stride)
        return torch.nn.functional.conv2d(x, self.weight, self.bias, padding=1)


SW_CFG = dict(classes=5, base_size=64, crop_h=33, crop_w=33, scales=[0.75, 1.0, 1.5],
              mean=[0.485 * 255, 0.456 * 255, 0.406 * 255], std=[0.229 * 255, 0.224 * 255, 0.225 * 255])


def sw_image(seed=9, h=40, w=60):
    rng = np.random.default_rng(seed)
    return (rng.random((h, w, 3)) * 255).astype(np.float32)


def metric_case(seed, shape, K, ignore=255):
    """Synthetic (prediction, target) pair of the metric goldens (identical to tests/golden/_ref_worker.py)."""
    rng = np.random.default_rng(seed)
    target = rng.integers(0, K, size=shape).astype(np.int64)
    pred = np.where(rng.random(shape) < 0.6, target, rng.integers(0, K, size=shape)).astype(np.int64)
    target[rng.random(shape) < 0.07] = ignore
    return pred, target


METRIC_CASES = [(1, (2, 33, 47), 150), (2, (1, 65, 65), 19), (3, (4000,), 2), (4, (3, 17), 300)]
