"""Oracle: evaluation tables (numpy).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows `cellulus/evaluate.py:72-105`; pinned by
`tests/golden/evaluate.npz` (outputs of the reference's own `compute_pairwise_IoU` / `compute_F1`).
"""

from __future__ import annotations

import numpy as np


def compute_pairwise_IoU(prediction: np.ndarray, groundtruth: np.ndarray):
    """`evaluate.py:72-98`: one full-image comparison per (prediction id, ground-truth id) pair."""
    prediction_ids = np.unique(prediction)
    prediction_ids = prediction_ids[prediction_ids != 0]
    groundtruth_ids = np.unique(groundtruth)
    groundtruth_ids = groundtruth_ids[groundtruth_ids != 0]
    if len(groundtruth_ids) == 0:
        return None
    iou = np.zeros((len(prediction_ids), len(groundtruth_ids)), dtype=float)
    iog = np.zeros_like(iou)
    for j, p in enumerate(prediction_ids):
        in_p = prediction == p
        for k, g in enumerate(groundtruth_ids):
            in_g = groundtruth == g
            inter = np.sum(in_p & in_g)
            iou[j, k] = inter / np.sum(in_p | in_g)
            iog[j, k] = inter / np.sum(in_g)
    return iou, np.sum(iou[iog > 0.5]), len(groundtruth_ids)


def compute_F1(IoU_table: np.ndarray, threshold=0.5):
    """`evaluate.py:101-105`."""
    hit = IoU_table > threshold
    FP = np.sum(np.sum(hit, axis=1) == 0)
    FN = np.sum(np.sum(hit, axis=0) == 0)
    TP = IoU_table.shape[1] - FN
    return 2 * TP / (2 * TP + FP + FN), TP, FP, FN
