"""Oracle: test-time-augmentation aggregate (torch CPU).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Follows
`cellulus/models/unet.py:90-98`: population std / mean over the T noisy
predictions, per-channel std summed over channels, `cat((mean, std))`.
"""

from __future__ import annotations

import torch


def tta_aggregate(predictions: torch.Tensor) -> torch.Tensor:
    """`predictions` (T, C, *S) fp32 -> (C+1, *S) fp32 (`models/unet.py:90-98`)."""
    embedding_std, embedding_mean = torch.std_mean(
        predictions, dim=0, keepdim=False, unbiased=False
    )
    embedding_std = embedding_std.sum(dim=0, keepdim=True)
    return torch.cat((embedding_mean, embedding_std), dim=0)


def tta_aggregate_float64(predictions: torch.Tensor) -> torch.Tensor:
    """Exact-arithmetic yardstick for the fp32 aggregate (two-pass, float64)."""
    return tta_aggregate(predictions.double())
