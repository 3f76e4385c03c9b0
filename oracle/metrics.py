"""ORACLE — TEST INFRASTRUCTURE ONLY. Restatement of the reference's per-class area counters
(util/util.py:40-52 numpy version, :55-67 torch version): after forcing the prediction to `ignore_index` wherever the
target is ignored, count per class k in [0, K): pixels where prediction == target == k, pixels predicted k, pixels
labelled k; union = predicted + labelled - intersection. Pinned against the reference's own numpy function in
tests/golden/metrics.npz (tests/golden/_ref_worker.py::golden_metrics)."""
import numpy as np


def intersection_and_union(pred, target, K, ignore_index=255):
    pred = np.asarray(pred).reshape(-1).astype(np.int64).copy()
    target = np.asarray(target).reshape(-1).astype(np.int64)
    pred[target == ignore_index] = ignore_index

    def count(v):
        v = v[(v >= 0) & (v < K)]
        return np.bincount(v, minlength=K).astype(np.int64)

    inter = count(pred[pred == target])
    area_pred, area_target = count(pred), count(target)
    return inter, area_pred + area_target - inter, area_target, pred
