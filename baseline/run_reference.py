"""Times the reference's OWN modules (baseline/_ref/model/pspnet.py, model/psanet.py) on a synthetic training step.

Run by bench.py in a subprocess with cwd = PYTHONPATH = baseline/_ref (the reference's `model` / `lib` packages must not
meet this repository's packages of the same name). Mirrors tool/train.py: criterion and model construction :121-132, the
eight SGD parameter groups :125-140, nn.DataParallel(model.cuda()) for the single-process GPU case :159, and the step
body :267-276. Prints one JSON object.

    python run_reference.py --device cpu|cuda --arch psp --layers 50 --classes 150 --size 473 --batch 2 \
        --steps 4 --warmup 1 [--threads 32]
"""
import argparse
import json
import os
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--arch", default="psp")
    ap.add_argument("--layers", type=int, default=50)
    ap.add_argument("--classes", type=int, default=150)
    ap.add_argument("--size", type=int, default=473)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()

    import torch
    import torch.nn as nn
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    import model.pspnet as ref_pspnet          # the reference's own file (cwd / PYTHONPATH = baseline/_ref)
    assert os.path.abspath(ref_pspnet.__file__).startswith(os.path.abspath(os.getcwd())), ref_pspnet.__file__
    torch.manual_seed(0)
    criterion = nn.CrossEntropyLoss(ignore_index=255)
    if a.arch == "psp":
        model = ref_pspnet.PSPNet(layers=a.layers, classes=a.classes, zoom_factor=8, criterion=criterion,
                                  pretrained=False)
        new = [model.ppm, model.cls, model.aux]
    else:
        from model.psanet import PSANet
        mk = 2 * ((a.size - 1) // 16 + 1) - 1
        model = PSANet(layers=a.layers, classes=a.classes, zoom_factor=8, psa_type=2, compact=False, shrink_factor=2,
                       mask_h=mk, mask_w=mk, normalization_factor=1.0, psa_softmax=True, criterion=criterion,
                       pretrained=False)
        new = [model.psa, model.cls, model.aux]
    ori = [model.layer0, model.layer1, model.layer2, model.layer3, model.layer4]
    groups = [dict(params=m.parameters(), lr=0.01) for m in ori] + [dict(params=m.parameters(), lr=0.1) for m in new]
    opt = torch.optim.SGD(groups, lr=0.01, momentum=0.9, weight_decay=1e-4)
    dev = torch.device(a.device)
    if dev.type == "cuda":
        model = torch.nn.DataParallel(model.cuda())     # tool/train.py:159 (single process, non-distributed)
    model.train()

    g = torch.Generator().manual_seed(a.seed)
    x = torch.randn((a.batch, 3, a.size, a.size), generator=g)
    y = torch.randint(0, a.classes, (a.batch, a.size, a.size), generator=g)
    y[torch.rand((a.batch, a.size, a.size), generator=g) < 0.05] = 255
    if dev.type == "cuda":
        x, y = x.pin_memory(), y.pin_memory()

    def step():
        inp, tgt = (x.cuda(non_blocking=True), y.cuda(non_blocking=True)) if dev.type == "cuda" else (x, y)
        _, main_loss, aux_loss = model(inp, tgt)
        main_loss, aux_loss = torch.mean(main_loss), torch.mean(aux_loss)      # tool/train.py:270-271
        loss = main_loss + 0.4 * aux_loss
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        seconds = e0.elapsed_time(e1) / 1e3
    else:
        t0 = time.perf_counter()
        for _ in range(a.steps):
            loss = step()
        seconds = time.perf_counter() - t0
    print(json.dumps({"images_per_sec": a.batch * a.steps / seconds, "seconds": seconds, "steps": a.steps,
                      "batch": a.batch, "threads": torch.get_num_threads(), "device": a.device,
                      "loss": float(loss.item()), "module_file": os.path.relpath(ref_pspnet.__file__),
                      "tf32_conv": bool(torch.backends.cudnn.allow_tf32), "torch": torch.__version__}))


if __name__ == "__main__":
    sys.exit(main())
