"""Loss factory with the reference's call signature (`cellulus/criterions/__init__.py:4-17`)."""

from cellulus_b200.criterions.oce_loss import (  # noqa: F401
    GraphedLossCycle,
    GraphedLossStep,
    OCELoss,
    oce_loss_fused,
    oce_loss_fused_sampled,
)


def get_loss(temperature, regularizer_weight, density, num_spatial_dims, device):
    """`OCELoss` module; its `.fused(offsets, anchors, refs)` is the one-kernel form of the loss slice."""
    settings = (temperature, regularizer_weight, density, num_spatial_dims, device)
    return OCELoss(*settings)
