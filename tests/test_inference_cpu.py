"""CPU tier for the sliding-window inference row: the serial oracle (oracle/sliding_window.py) is pinned
against outputs of the reference's own tool/test.py functions (tests/golden/sliding_window.npz), and the batched engine
(semseg_b200/inference.py — device-agnostic torch glue) is checked against the oracle with a tiny CPU network."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import sliding_window as osw
from tests import util


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "sliding_window.npz"))


@pytest.mark.parametrize("tag,stride", [("s1", 1), ("s8", 8)])
def test_oracle_matches_reference_goldens(golden, tag, stride):
    torch.set_num_threads(8)
    c = util.SW_CFG
    image = util.sw_image()
    assert np.array_equal(image, golden["image"])
    h, w, _ = image.shape
    model = util.TinySegNet(c["classes"], stride).eval()
    # same torch, same thread count, same one-crop-at-a-time calls as the reference -> bit-identical
    assert np.array_equal(osw.score_crop(model, image[:33, :33].copy(), c["mean"], c["std"]), golden[tag + "/net_process"])
    assert np.array_equal(osw.score_crop(model, image[3:36, 7:40].copy(), c["mean"], None, flip=False),
                          golden[tag + "/net_process_nostd_noflip"])
    sp = osw.score_scale(model, image, c["classes"], c["crop_h"], c["crop_w"], h, w, c["mean"], c["std"])
    assert np.array_equal(sp, golden[tag + "/scale_process"])
    scores, amax = osw.score_image(model, image, c["classes"], c["mean"], c["std"], c["base_size"], c["crop_h"],
                                   c["crop_w"], c["scales"])
    assert np.array_equal(scores, golden[tag + "/scores"])
    assert np.array_equal(amax, golden[tag + "/argmax"])


@pytest.mark.parametrize("stride", [1, 8])
@pytest.mark.parametrize("max_batch", [2, 6, 64])
def test_batched_engine_matches_serial_oracle(golden, stride, max_batch):
    from semseg_b200 import inference
    torch.set_num_threads(8)
    c = util.SW_CFG
    image = util.sw_image()
    h, w, _ = image.shape
    model = util.TinySegNet(c["classes"], stride).eval()
    eng = inference.SlidingWindowPredictor(model, c["classes"], c["crop_h"], c["crop_w"], c["mean"], c["std"],
                                           max_batch=max_batch)
    scores, amax = eng(image, c["base_size"], c["scales"], exact=True)
    ref_scores, ref_amax = osw.score_image(model, image, c["classes"], c["mean"], c["std"], c["base_size"],
                                           c["crop_h"], c["crop_w"], c["scales"])
    # a CPU conv may pick another blocking for another batch size: allow fp32 rounding, not more
    assert np.allclose(scores, ref_scores, rtol=0, atol=2e-6)
    assert (amax != ref_amax).mean() < 1e-3
    assert np.allclose(scores, golden["s%d/scores" % stride], rtol=0, atol=2e-6)
    n_crops = 0
    for s in c["scales"]:
        nh, nw = inference.scaled_size(h, w, round(s * c["base_size"]))
        n_crops += len(inference.crop_origins(max(nh, c["crop_h"]), c["crop_h"])) * \
            len(inference.crop_origins(max(nw, c["crop_w"]), c["crop_w"]))
    assert eng.forward_calls == sum(
        math.ceil(len(inference.crop_origins(max(inference.scaled_size(h, w, round(s * c["base_size"]))[0], c["crop_h"]),
                                             c["crop_h"])) *
                  len(inference.crop_origins(max(inference.scaled_size(h, w, round(s * c["base_size"]))[1], c["crop_w"]),
                                             c["crop_w"])) / (max_batch // 2)) for s in c["scales"])
    assert eng.forward_calls <= n_crops          # the reference makes one model call per crop
    # device-side resize / accumulation / argmax: same sampling as cv2.INTER_LINEAR, weights in fp64 instead of fp32
    fast_scores, fast_amax = eng(image, c["base_size"], c["scales"])
    assert np.allclose(fast_scores, ref_scores, rtol=0, atol=5e-6)
    assert (fast_amax != ref_amax).mean() < 1e-3
    none_scores, amax2 = eng(image, c["base_size"], c["scales"], return_scores=False)
    assert none_scores is None and np.array_equal(amax2, fast_amax)


def test_reference_named_entry_points(golden):
    from semseg_b200 import inference
    c = util.SW_CFG
    image = util.sw_image()
    h, w, _ = image.shape
    model = util.TinySegNet(c["classes"], 8).eval()
    a = inference.net_process(model, image[:33, :33].copy(), c["mean"], c["std"])
    assert a.dtype == np.float32 and np.allclose(a, golden["s8/net_process"], rtol=0, atol=2e-6)
    b = inference.net_process(model, image[3:36, 7:40].copy(), c["mean"], None, flip=False)
    assert np.allclose(b, golden["s8/net_process_nostd_noflip"], rtol=0, atol=2e-6)
    sp = inference.scale_process(model, image, c["classes"], c["crop_h"], c["crop_w"], h, w, c["mean"], c["std"])
    assert sp.dtype == np.float64 and sp.shape == (h, w, c["classes"])
    assert np.allclose(sp, golden["s8/scale_process"], rtol=0, atol=2e-6)


def test_crop_grid_properties():
    from semseg_b200 import inference
    for crop in (33, 65, 473, 713):
        for extent in (crop, crop + 1, crop + 100, 2 * crop, 1024, 2048, 3 * crop + 7):
            if extent < crop:
                continue
            o = inference.crop_origins(extent, crop)
            assert o == osw.crop_origins(extent, crop)
            stride = int(math.ceil(crop * 2 / 3))
            assert len(o) == int(math.ceil(float(extent - crop) / stride) + 1)      # tool/test.py:160-161
            assert o[0] == 0 and o[-1] == extent - crop and all(0 <= v <= extent - crop for v in o)
            covered = np.zeros(extent, dtype=bool)
            for v in o:
                covered[v:v + crop] = True
            assert covered.all()
    assert inference.scaled_size(1024, 2048, 2048) == (1024, 2048)
    assert inference.scaled_size(806, 512, 512) == (512, round(512 / 806.0 * 512))
