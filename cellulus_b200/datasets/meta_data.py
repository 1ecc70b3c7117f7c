"""`DatasetMetaData` (`cellulus/datasets/meta_data.py`): the zarr layout contract of the reference -- arrays
are `(s, c, [t,] [z,] y, x)` and name their axes in an `axis_names` attribute."""

from __future__ import annotations

from typing import Tuple

from cellulus_b200 import zarr_lite

_LAYOUT = ('expected layout: (s, c, [t,] [z,] y, x) = samples, channels, frames, spatial extents, with an '
           '"axis_names" attribute such as ["s", "c", "y", "x"]')


def _fail(what: str):
    raise RuntimeError(f"{what}\n\n{_LAYOUT}")


class DatasetMetaData:
    """Axis bookkeeping: which axis holds samples / channels / time, and the spatial extents in array order."""

    def __init__(self, shape, axis_names):
        position = {name: dim for dim, name in enumerate(axis_names)}
        self.num_dims = len(axis_names)
        self.sample_dim = position.get("s")
        self.channel_dim = position.get("c")
        self.time_dim = position.get("t")
        self.num_samples: int = shape[self.sample_dim] if self.sample_dim is not None else 0
        self.num_channels: int = shape[self.channel_dim] if self.channel_dim is not None else 0
        spatial = [dim for dim, name in enumerate(axis_names) if name in ("z", "y", "x")]
        # 't' counts as a spatial dimension but contributes no extent (reference quirk Q14)
        self.num_spatial_dims: int = len(spatial) + (1 if self.time_dim is not None else 0)
        self.spatial_array: Tuple[int, ...] = tuple(shape[dim] for dim in spatial)
        if self.sample_dim is None:
            _fail("dataset does not have a sample dimension")
        if self.channel_dim is None:
            _fail("dataset does not have a channel dimension")
        if self.num_dims != len(shape):
            _fail(f"dataset has {len(shape)} dimensions, but attribute axis_names has {self.num_dims} entries")

    @staticmethod
    def from_dataset_config(dataset_config) -> "DatasetMetaData":
        where = f'"{dataset_config.dataset_name}" in {dataset_config.container_path}'
        container = zarr_lite.open(dataset_config.container_path, "r")
        try:
            data = container[dataset_config.dataset_name]
        except KeyError:
            _fail(f"no dataset {where}")
        try:
            axis_names = data.attrs["axis_names"]
        except KeyError:
            _fail(f'dataset {where} has no "axis_names" attribute')
        try:
            return DatasetMetaData(data.shape, axis_names)
        except RuntimeError as error:
            raise RuntimeError(f"dataset {where} has invalid meta-data") from error
