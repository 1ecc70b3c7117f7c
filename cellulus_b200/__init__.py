"""cellulus_b200 -- B200-native (sm_100a) embedding-space hot path of funkelab/cellulus.

Same call signatures as the reference for the functions on that path:

    cellulus_b200.criterions.get_loss / OCELoss          (cellulus/criterions)
    cellulus_b200.criterions.oce_loss_fused              (fused gather + loss + backward)
    cellulus_b200.models.UNetModel.select_and_add_coordinates, tta_aggregate
    cellulus_b200.utils.mean_shift.mean_shift_segmentation
    cellulus_b200.utils.misc.size_filter
    cellulus_b200.detect.detect_embeddings / threshold_otsu

All compute is hand-written CUDA behind the C ABI in include/cellulus_b200.h
(libcellulus_b200.so, built by `python -m cellulus_b200.build`).  There is no
CPU fallback: without the library, or with CPU tensors, the ops raise.
"""

__version__ = "0.1.0"
