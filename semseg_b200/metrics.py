"""Device-side evaluation metrics of the training / validation loop.

`intersectionAndUnionGPU` has the signature, return values and in-place side effect of the reference's helper
(util/util.py:55-67, used at tool/train.py:286 and :375) but is ONE kernel with integer shared-memory histograms
instead of a masked write, a boolean-mask gather and three `torch.histc` passes. Replace the import in tool/train.py:21
by `from semseg_b200.metrics import intersectionAndUnionGPU`.
"""
import ctypes

import torch

from . import _lib

__all__ = ["intersectionAndUnionGPU"]


def intersectionAndUnionGPU(output, target, K, ignore_index=255):
    """output, target: int64 CUDA tensors of the same shape (N, N*L or N*H*W), class ids in [0, K).
    Returns (area_intersection, area_union, area_target): float32 [K] tensors on the device. Like the reference,
    `output` is overwritten with `ignore_index` wherever `target == ignore_index`."""
    assert output.dim() in [1, 2, 3]
    assert output.shape == target.shape
    if not (output.is_cuda and target.is_cuda):
        raise _lib.SemsegError("intersectionAndUnionGPU needs CUDA tensors (no CPU fallback); use the reference's "
                               "numpy intersectionAndUnion on the host")
    assert output.dtype == torch.int64 and target.dtype == torch.int64
    lib = _lib.load()
    out_flat = output.view(-1)                  # a view: the masking below is visible to the caller, as in the reference
    tgt_flat = target.reshape(-1)
    counts = torch.empty((3, int(K)), dtype=torch.int32, device=output.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(output.device).cuda_stream)
    _lib.check(lib.semseg_iou_hist(ctypes.c_void_p(out_flat.data_ptr()), ctypes.c_void_p(tgt_flat.data_ptr()),
                                   out_flat.numel(), int(K), int(ignore_index), 1,
                                   ctypes.c_void_p(counts.data_ptr()), stream), "semseg_iou_hist")
    c = counts.float()
    return c[0], c[1] + c[2] - c[0], c[2]
