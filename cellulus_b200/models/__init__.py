"""Model factory with the reference's call signature (`cellulus/models/__init__.py:6-23`)."""

from cellulus_b200.models.unet import TTAAccumulator, UNetModel, tta_aggregate  # noqa: F401


def get_model(in_channels, out_channels, num_fmaps, fmap_inc_factor, features_in_last_layer, downsampling_factors,
              num_spatial_dims) -> UNetModel:
    """U-Net stand-in + 1x1 head; every argument is handed to `UNetModel` by name."""
    arguments = dict(locals())
    return UNetModel(**arguments)
