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
