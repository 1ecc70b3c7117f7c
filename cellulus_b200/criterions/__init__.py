from cellulus_b200.criterions.oce_loss import GraphedLossStep, OCELoss, oce_loss_fused  # noqa: F401


def get_loss(
    temperature,
    regularizer_weight,
    density,
    num_spatial_dims,
    device,
):
    """Same factory as `cellulus/criterions/__init__.py:4-17`."""
    return OCELoss(
        temperature,
        regularizer_weight,
        density,
        num_spatial_dims,
        device,
    )
