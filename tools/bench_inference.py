"""Multi-scale sliding-window inference throughput (SURVEY §8 f3 / BASELINE config 5 shape): one synthetic
1024x2048 image, scales {0.5 ... 1.75}, base size 2048, crop 713 (PSPNet101, 19 classes) or a lighter PSPNet50 / 473
setting. Three arms through the SAME eval-mode network of this package:
  reference_procedure_serial : one crop + its mirror per model call and the host-side cv2 / numpy finish of
                               tool/test.py:122-199 (the reference's procedure, driven by this package's network)
  reference_finish_batched   : crops batched (`--max-batch`), host-side finish (bit-identical scores)
  device_finish_batched      : crops batched, resize / sum over scales / argmax on the device (the default)
Prints one JSON line with images/s of the arms. Not part of bench.py's contract (that one measures the training step).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from semseg_b200 import inference  # noqa: E402
from semseg_b200.pspnet import PSPNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=101)
    ap.add_argument("--classes", type=int, default=19)
    ap.add_argument("--crop", type=int, default=713)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--base-size", type=int, default=2048)
    ap.add_argument("--scales", type=float, nargs="+", default=[0.5, 0.75, 1.0, 1.25, 1.5, 1.75])
    ap.add_argument("--max-batch", type=int, default=32)
    ap.add_argument("--repeats", type=int, default=1)
    args = ap.parse_args()
    torch.manual_seed(0)
    model = PSPNet(layers=args.layers, classes=args.classes, zoom_factor=8, pretrained=False).cuda().eval()
    rng = np.random.default_rng(0)
    image = (rng.random((args.height, args.width, 3)) * 255).astype(np.float32)
    mean = [0.485 * 255, 0.456 * 255, 0.406 * 255]
    std = [0.229 * 255, 0.224 * 255, 0.225 * 255]
    out = {"workload": "PSPNet%d eval, %dx%d image, crop %d, scales %s, flip" % (args.layers, args.height, args.width,
                                                                              args.crop, args.scales)}
    results = {}
    arms = (("device_finish_batched", args.max_batch, False),      # this package's default
            ("reference_finish_batched", args.max_batch, True),    # batched network calls, host cv2 / numpy finish
            ("reference_procedure_serial", 2, True))               # one crop (+ mirror) per call, host finish
    for name, mb, exact in arms:
        eng = inference.SlidingWindowPredictor(model, args.classes, args.crop, args.crop, mean, std, max_batch=mb)
        eng(image, args.base_size, args.scales[:1], exact=exact, return_scores=False)   # warm-up (packing, allocator)
        torch.cuda.synchronize()
        eng.forward_calls = 0
        t0 = time.perf_counter()
        for _ in range(args.repeats):
            _, amax = eng(image, args.base_size, args.scales, exact=exact, return_scores=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.repeats
        results[name] = amax
        out[name] = {"seconds_per_image": dt, "images_per_sec": 1.0 / dt,
                     "model_calls_per_image": eng.forward_calls // args.repeats, "max_batch": mb}
    ref = results["reference_procedure_serial"]
    out["argmax_identical_reference_finish"] = bool(np.array_equal(results["reference_finish_batched"], ref))
    out["argmax_mismatch_device_finish"] = float((results["device_finish_batched"] != ref).mean())
    out["speedup_vs_reference_procedure"] = (out["reference_procedure_serial"]["seconds_per_image"] /
                                             out["device_finish_batched"]["seconds_per_image"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
