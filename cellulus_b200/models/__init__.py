from cellulus_b200.models.unet import TTAAccumulator, UNetModel, tta_aggregate  # noqa: F401
