"""The reference's TOML config schema (`cellulus/configs/*.py`), kept field for field so that existing
experiment files load unchanged: `ExperimentConfig(**tomllib.load(f))`.

One module instead of six; same class names, field names, defaults and type checks
(`configs/dataset_config.py:37-41`, `model_config.py:50-59`, `train_config.py:104-127`,
`inference_config.py:119-159`, `experiment_config.py:48-62`).  Nested tables become nested config
objects (`configs/utils.py:4-11`).
"""

from __future__ import annotations

from datetime import datetime
from pathlib import Path
from typing import List

import attrs
from attrs.validators import in_, instance_of, optional


def _nested(cls):
    """dict -> cls(**dict); None and ready-made instances pass through."""

    def convert(value):
        if value is None or isinstance(value, cls):
            return value
        return cls(**value)

    return convert


def _maybe_path(value):
    return None if value is None else Path(value)


@attrs.define
class DatasetConfig:
    """A zarr container + dataset name (+ the dataset a stage reads its input from)."""

    container_path: Path = attrs.field(converter=Path)
    dataset_name: str = attrs.field(validator=instance_of(str))
    secondary_dataset_name: str = attrs.field(default=None, validator=optional(instance_of(str)))


@attrs.define
class ModelConfig:
    num_fmaps: int = attrs.field(validator=instance_of(int))
    fmap_inc_factor: int = attrs.field(validator=instance_of(int))
    features_in_last_layer: int = attrs.field(default=64)
    downsampling_factors: List[List[int]] = attrs.field(default=[[2, 2]])
    checkpoint: Path = attrs.field(default=None, converter=_maybe_path)
    initialize: bool = attrs.field(default=True, validator=instance_of(bool))


@attrs.define
class TrainConfig:
    train_data_config: DatasetConfig = attrs.field(default=None, converter=_nested(DatasetConfig))
    validate_data_config: DatasetConfig = attrs.field(default=None, converter=_nested(DatasetConfig))
    crop_size: List = attrs.field(default=[252, 252], validator=instance_of(List))
    batch_size: int = attrs.field(default=8, validator=instance_of(int))
    max_iterations: int = attrs.field(default=100_000, validator=instance_of(int))
    initial_learning_rate: float = attrs.field(default=4e-5, validator=instance_of(float))
    density: float = attrs.field(default=0.1, validator=instance_of(float))
    kappa: float = attrs.field(default=10.0, validator=instance_of(float))
    temperature: float = attrs.field(default=10.0, validator=instance_of(float))
    regularizer_weight: float = attrs.field(default=1e-5, validator=instance_of(float))
    save_model_every: int = attrs.field(default=1_000, validator=instance_of(int))
    save_best_model_every: int = attrs.field(default=100, validator=instance_of(int))
    save_snapshot_every: int = attrs.field(default=1_000, validator=instance_of(int))
    num_workers: int = attrs.field(default=8, validator=instance_of(int))
    elastic_deform: bool = attrs.field(default=True, validator=instance_of(bool))
    control_point_spacing: int = attrs.field(default=64, validator=instance_of(int))
    control_point_jitter: float = attrs.field(default=2.0, validator=instance_of(float))
    device: str = attrs.field(default="cuda:0", validator=instance_of(str))


@attrs.define
class InferenceConfig:
    dataset_config: DatasetConfig = attrs.field(default=None, converter=_nested(DatasetConfig))
    prediction_dataset_config: DatasetConfig = attrs.field(default=None, converter=_nested(DatasetConfig))
    detection_dataset_config: DatasetConfig = attrs.field(default=None, converter=_nested(DatasetConfig))
    segmentation_dataset_config: DatasetConfig = attrs.field(default=None, converter=_nested(DatasetConfig))
    evaluation_dataset_config: DatasetConfig = attrs.field(default=None, converter=_nested(DatasetConfig))
    device: str = attrs.field(default="cuda:0", validator=instance_of(str))
    crop_size: List = attrs.field(default=[252, 252], validator=instance_of(List))
    p_salt_pepper = attrs.field(default=0.01, validator=instance_of(float))
    num_infer_iterations = attrs.field(default=16, validator=instance_of(int))
    threshold = attrs.field(default=None, validator=optional(instance_of(float)))
    clustering = attrs.field(default="meanshift", validator=in_(["meanshift", "greedy"]))
    use_seeds = attrs.field(default=False, validator=instance_of(bool))
    bandwidth = attrs.field(default=None, validator=optional(instance_of(float)))
    num_bandwidths = attrs.field(default=1, validator=instance_of(int))
    reduction_probability = attrs.field(default=0.1, validator=instance_of(float))
    min_size = attrs.field(default=None, validator=optional(instance_of(int)))
    post_processing = attrs.field(default="cell", validator=in_(["cell", "nucleus"]))
    grow_distance = attrs.field(default=3, validator=instance_of(int))
    shrink_distance = attrs.field(default=6, validator=instance_of(int))


@attrs.define
class ExperimentConfig:
    model_config: ModelConfig = attrs.field(converter=_nested(ModelConfig))
    experiment_name: str = attrs.field(default=datetime.today().strftime("%Y-%m-%d"), validator=instance_of(str))
    normalization_factor: float = attrs.field(default=None, validator=optional(instance_of(float)))
    object_size: int = attrs.field(default=30, validator=instance_of(int))
    train_config: TrainConfig = attrs.field(default=None, converter=_nested(TrainConfig))
    inference_config: InferenceConfig = attrs.field(default=None, converter=_nested(InferenceConfig))


__all__ = ["DatasetConfig", "ModelConfig", "TrainConfig", "InferenceConfig", "ExperimentConfig"]
