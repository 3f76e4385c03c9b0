"""Runs INSIDE the reference tree (cwd = reference copy, PYTHONPATH = reference copy only). See make_golden.py."""
import hashlib
import json
import os

import numpy as np
import torch

OUT = os.environ["GOLDEN_OUT"]
torch.set_num_threads(8)


def synth(n, h, w, classes, seed=123):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, 3, h, w), generator=g)
    y = torch.randint(0, classes, (n, h, w), generator=g)
    ign = torch.rand((n, h, w), generator=g) < 0.05
    y[ign] = 255
    return x, y


def checksum(sd):
    return {k: [float(v.double().abs().sum()), float(v.double().sum())] for k, v in sd.items()
            if v.dtype.is_floating_point}


GRAD_KEYS_COMMON = ["layer0.0.weight", "layer0.1.weight", "layer1.0.conv1.weight", "layer2.0.conv2.weight",
                    "layer2.0.downsample.0.weight", "layer3.0.conv2.weight", "layer3.5.bn2.weight",
                    "layer4.2.conv3.weight", "layer4.0.bn3.bias", "cls.0.weight", "cls.1.weight", "cls.4.weight",
                    "cls.4.bias", "aux.0.weight", "aux.4.bias"]


def run_model(model, x, y, tag, extra_grad_keys):
    model.train()
    out, main_loss, aux_loss = model(x, y)
    loss = main_loss + 0.4 * aux_loss
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    res = {
        "main_loss": np.float64(main_loss.item()), "aux_loss": np.float64(aux_loss.item()),
        "argmax": out.numpy().astype(np.int16),
    }
    for k in GRAD_KEYS_COMMON + extra_grad_keys:
        res["gradnorm/" + k] = np.float64(grads[k].double().norm().item())
        res["gradhead/" + k] = grads[k].flatten()[:16].numpy().copy()
    res["gradnorm_total"] = np.float64(torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item())
    model.eval()
    with torch.no_grad():
        logits = model(x)
    res["eval_logits_s8"] = logits[:, :, ::8, ::8].numpy().copy()
    res["eval_logits_absmean"] = np.float64(logits.abs().mean().item())
    res["running_mean/layer4.2.bn3"] = dict(model.named_buffers())["layer4.2.bn3.running_mean"].numpy()[:32].copy()
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **res)
    print(tag, "main", main_loss.item(), "aux", aux_loss.item())


def golden_psamask():
    import lib.psa.functional as PF
    rng = np.random.default_rng(7)
    res = {}
    for (n, h, w, mh, mw) in [(2, 4, 5, 7, 9), (1, 6, 7, 5, 3), (2, 5, 5, 9, 9), (1, 30, 30, 59, 59)]:
        for t in (0, 1):
            key = "n%d_h%d_w%d_mh%d_mw%d_t%d" % (n, h, w, mh, mw, t)
            x = rng.standard_normal((n, mh * mw, h, w)).astype(np.float32)
            xt = torch.from_numpy(x).requires_grad_(True)
            o = PF.psa_mask(xt, t, mh, mw)
            g = rng.standard_normal(tuple(o.shape)).astype(np.float32)
            o.backward(torch.from_numpy(g))
            on, gn = o.detach().numpy(), xt.grad.numpy()
            res[key + "/out_sha"] = hashlib.sha256(on.tobytes()).hexdigest()
            res[key + "/din_sha"] = hashlib.sha256(gn.tobytes()).hexdigest()
            if on.size <= 20000:
                res[key + "/out"] = on
                res[key + "/din"] = gn
    np.savez_compressed(os.path.join(OUT, "psamask.npz"), **res)
    print("psamask goldens written")


class TinySegNet(torch.nn.Module):
    """Deterministic stand-in network for the sliding-window goldens (weights from a seeded generator, independent of
    nn init order). stride = 1: logits at the input size; stride = 8: logits at 1/8 size (the caller resizes)."""

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


def golden_sliding_window():
    """Outputs of the reference's own net_process / scale_process and of its evaluation-loop body (tool/test.py:122-199)
    for two tiny deterministic networks. `.cuda()` is made a no-op (this container has no GPU); nothing else is shimmed."""
    import importlib.util
    import cv2
    spec = importlib.util.spec_from_file_location("ref_tool_test", os.path.join(os.getcwd(), "tool", "test.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.Tensor.cuda = lambda self, *a, **k: self
    c = SW_CFG
    image = sw_image()
    h, w, _ = image.shape
    res = {"image": image}
    for tag, stride in (("s1", 1), ("s8", 8)):
        model = TinySegNet(c["classes"], stride).eval()
        res[tag + "/net_process"] = ref.net_process(model, image[:33, :33].copy(), c["mean"], c["std"])
        res[tag + "/net_process_nostd_noflip"] = ref.net_process(model, image[3:36, 7:40].copy(), c["mean"], None, flip=False)
        res[tag + "/scale_process"] = ref.scale_process(model, image, c["classes"], c["crop_h"], c["crop_w"], h, w,
                                                        c["mean"], c["std"])
        # body of the evaluation loop, tool/test.py:186-199
        prediction = np.zeros((h, w, c["classes"]), dtype=float)
        for scale in c["scales"]:
            long_size = round(scale * c["base_size"])
            new_h = long_size
            new_w = long_size
            if h > w:
                new_w = round(long_size / float(h) * w)
            else:
                new_h = round(long_size / float(w) * h)
            image_scale = cv2.resize(image, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
            prediction += ref.scale_process(model, image_scale, c["classes"], c["crop_h"], c["crop_w"], h, w,
                                            c["mean"], c["std"])
        prediction /= len(c["scales"])
        res[tag + "/scores"] = prediction
        res[tag + "/argmax"] = np.argmax(prediction, axis=2).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "sliding_window.npz"), **res)
    print("sliding-window goldens written")


def metric_case(seed, shape, K, ignore=255):
    rng = np.random.default_rng(seed)
    target = rng.integers(0, K, size=shape).astype(np.int64)
    pred = np.where(rng.random(shape) < 0.6, target, rng.integers(0, K, size=shape)).astype(np.int64)
    target[rng.random(shape) < 0.07] = ignore
    return pred, target


METRIC_CASES = [(1, (2, 33, 47), 150), (2, (1, 65, 65), 19), (3, (4000,), 2), (4, (3, 17), 300)]


def golden_metrics():
    """Outputs of the reference's numpy intersectionAndUnion (util/util.py:40-52; the torch twin at :55-67 is the same
    arithmetic but needs CUDA for histc on int64)."""
    from util.util import intersectionAndUnion
    res = {}
    for seed, shape, K in METRIC_CASES:
        pred, target = metric_case(seed, shape, K)
        i, u, t = intersectionAndUnion(pred.copy(), target.copy(), K, 255)
        key = "s%d" % seed
        res[key + "/i"], res[key + "/u"], res[key + "/t"] = i.astype(np.int64), u.astype(np.int64), t.astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **res)
    print("metric goldens written")


def main():
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "metrics":
        golden_metrics()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sliding":
        golden_sliding_window()
        return
    golden_psamask()
    golden_sliding_window()
    golden_metrics()

    from model.pspnet import PSPNet
    from model.psanet import PSANet
    meta = {}

    torch.manual_seed(0)
    m = PSPNet(layers=50, classes=150, zoom_factor=8, dropout=0.0, pretrained=False)
    meta["pspnet50_keys"] = {k: list(v.shape) for k, v in m.state_dict().items()}
    meta["pspnet50_wsum"] = checksum({k: v for k, v in m.state_dict().items()
                                      if k in ("layer0.0.weight", "layer3.0.conv2.weight", "cls.0.weight",
                                               "cls.4.bias", "ppm.features.3.1.weight", "aux.4.weight")})
    x, y = synth(2, 65, 65, 150)
    run_model(m, x, y, "pspnet50_65", ["ppm.features.0.1.weight", "ppm.features.3.2.weight"])

    torch.manual_seed(0)
    m = PSANet(layers=50, classes=150, zoom_factor=8, dropout=0.0, psa_type=2, compact=False, shrink_factor=2,
               mask_h=9, mask_w=9, pretrained=False)
    meta["psanet50_keys"] = {k: list(v.shape) for k, v in m.state_dict().items()}
    meta["psanet50_wsum"] = checksum({k: v for k, v in m.state_dict().items()
                                      if k in ("layer0.0.weight", "psa.reduce.0.weight", "psa.attention_p.3.weight",
                                               "psa.proj.0.weight", "cls.0.weight")})
    x, y = synth(2, 65, 65, 150, seed=321)
    run_model(m, x, y, "psanet50_65", ["psa.reduce.0.weight", "psa.attention.3.weight", "psa.attention_p.3.weight",
                                       "psa.proj.1.weight"])

    torch.manual_seed(0)
    m = PSPNet(layers=101, classes=19, zoom_factor=8, dropout=0.0, pretrained=False)
    meta["pspnet101_keys"] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(OUT, "meta.json"), "w") as fh:
        json.dump(meta, fh)
    print("done")


if __name__ == "__main__":
    main()
