"""`evaluate(inference_config)` (`cellulus/evaluate.py:9-105`): F1 and SEG against ground truth.

The reference builds its IoU / IoG tables with four full-image comparisons per (prediction id, ground-truth
id) pair.  Here ONE device pass (`cb200_contingency`) counts the joint label histogram; every intersection is
an entry of that table, the areas are its row / column sums and `union = area_p + area_g - intersection`, all
exact integers, so the float tables -- and F1, SEG, TP, FP, FN -- are bit-identical to the reference's.
"""

from __future__ import annotations

import numpy as np
import torch

from cellulus_b200 import kernels as K
from cellulus_b200 import zarr_lite
from cellulus_b200.datasets.meta_data import DatasetMetaData


def _ranks(present: np.ndarray):
    """ids (ascending, background dropped: `evaluate.py:73-76`) and the value -> row/column table (0 = background)."""
    ids = np.nonzero(present)[0]
    ids = ids[ids != 0]
    rank = np.zeros(len(present), np.int32)
    rank[ids] = np.arange(1, len(ids) + 1, dtype=np.int32)
    return ids, rank


def compute_pairwise_IoU(prediction, groundtruth, device="cuda"):
    """Drop-in for `evaluate.py:72-98`: `(IoU_table, SEG_sum, n_groundtruth_ids)` or None without ground truth."""
    pred = torch.from_numpy(np.ascontiguousarray(prediction).astype(np.uint16)).to(device)
    gt = torch.from_numpy(np.ascontiguousarray(groundtruth).astype(np.uint16)).to(device)
    pred_ids, rank_p = _ranks(K.label_presence(pred, 65535).cpu().numpy())
    gt_ids, rank_g = _ranks(K.label_presence(gt, 65535).cpu().numpy())
    if len(gt_ids) == 0:
        return None
    rows, cols = len(pred_ids) + 1, len(gt_ids) + 1
    table = K.contingency(pred, gt, torch.from_numpy(rank_p).to(device), torch.from_numpy(rank_g).to(device),
                          rows, cols).cpu().numpy().astype(np.int64)
    area_p = table.sum(axis=1)[1:, None]
    area_g = table.sum(axis=0)[None, 1:]
    intersection = table[1:, 1:]
    IoU_table = intersection / (area_p + area_g - intersection)  # np.sum(intersection) / np.sum(union)
    IoG_table = intersection / area_g
    IoU_table = IoU_table.astype(float)
    # SEG counts a match when strictly more than half of the ground-truth object is covered (:96-98)
    return IoU_table, np.sum(IoU_table[IoG_table > 0.5]), len(gt_ids)


def compute_F1(IoU_table, threshold=0.5):
    """`evaluate.py:101-105`."""
    IoU_table_thresholded = IoU_table > threshold
    FP = np.sum(np.sum(IoU_table_thresholded, axis=1) == 0)
    FN = np.sum(np.sum(IoU_table_thresholded, axis=0) == 0)
    TP = IoU_table.shape[1] - FN
    return 2 * TP / (2 * TP + FP + FN), TP, FP, FN


def evaluate(inference_config) -> None:
    meta = DatasetMetaData.from_dataset_config(inference_config.dataset_config)
    cfg = inference_config.evaluation_dataset_config
    f = zarr_lite.open(cfg.container_path)
    ds_segmentation = f[cfg.secondary_dataset_name]
    ds_groundtruth = f[cfg.dataset_name]
    for bandwidth in range(inference_config.num_bandwidths):
        sample_list, F1_list, SEG_list, TP_list, FP_list, FN_list = [], [], [], [], [], []
        SEG_dataset, n_ids_dataset = 0, 0
        for sample in range(meta.num_samples):
            groundtruth = np.asarray(ds_groundtruth[sample, 0]).astype(np.uint16)
            prediction = np.asarray(ds_segmentation[sample, bandwidth]).astype(np.uint16)
            returned = compute_pairwise_IoU(prediction, groundtruth)
            if returned is None:
                continue
            IoU, SEG_image, n_GTids_image = returned
            F1_image, TP_image, FP_image, FN_image = compute_F1(IoU)
            F1_list.append(F1_image)
            SEG_list.append(SEG_image / n_GTids_image)
            SEG_dataset += SEG_image
            n_ids_dataset += n_GTids_image
            TP_list.append(TP_image)
            FP_list.append(FP_image)
            FN_list.append(FN_image)
            sample_list.append(sample)
            print(f"{sample}: F1={F1_image:.3f}, SEG={SEG_image / n_GTids_image:.3f}")
        F1_dataset = 2 * sum(TP_list) / (2 * sum(TP_list) + sum(FP_list) + sum(FN_list))
        print(f"F1 for dataset  is {F1_dataset:.05f}")
        print(f"SEG for dataset  is {SEG_dataset / n_ids_dataset:.05f}")
        with open(f"results_bandwidth-{bandwidth}.txt", "w") as out:  # same file, same format (:51-69)
            out.writelines("file index, F1, SEG, TP, FP, FN \n")
            out.writelines("+++++++++++++++++++++++++++++++++\n")
            for i in range(len(sample_list)):
                out.writelines(f"{sample_list[i]}, {F1_list[i]:.05f}, {SEG_list[i]:.05f}, {TP_list[i]},"
                               f" {FP_list[i]}, {FN_list[i]}\n")
            out.writelines("+++++++++++++++++++++++++++++++++\n")
            out.writelines(f"F1 for complete dataset is {F1_dataset:.05f} \n")
            out.writelines(f"SEG for complete dataset is {SEG_dataset / n_ids_dataset:.05f} \n")
