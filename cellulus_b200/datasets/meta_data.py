"""`DatasetMetaData` of `cellulus/datasets/meta_data.py`: the zarr layout contract -- arrays are
`(s, c, [t,] [z,] y, x)` and carry an `axis_names` attribute."""

from __future__ import annotations

from typing import Tuple

from cellulus_b200 import zarr_lite
from cellulus_b200.configs import DatasetConfig

_HELP = (
    "The raw dataset should have shape (s, c, [t,] [z,] y, x), where s = # of samples, c = # of channels, "
    "t = # of frames, and z/y/x are spatial extents. The dataset should have an \"axis_names\" attribute that "
    'contains the names of the used axes, e.g., ["s", "c", "y", "x"] for a 2D dataset.'
)


class DatasetMetaData:
    def __init__(self, shape, axis_names):
        self.num_dims = len(axis_names)
        self.num_spatial_dims: int = 0
        self.num_samples: int = 0
        self.num_channels: int = 0
        self.sample_dim = None
        self.channel_dim = None
        self.time_dim = None
        self.spatial_array: Tuple[int, ...] = ()
        for dim, name in enumerate(axis_names):
            if name == "s":
                self.sample_dim, self.num_samples = dim, shape[dim]
            elif name == "c":
                self.channel_dim, self.num_channels = dim, shape[dim]
            elif name == "t":  # counted as spatial, not appended to spatial_array (reference quirk Q14)
                self.num_spatial_dims += 1
                self.time_dim = dim
            elif name in ("z", "y", "x"):
                self.num_spatial_dims += 1
                self.spatial_array += (shape[dim],)
        if self.sample_dim is None:
            raise RuntimeError("dataset does not have a sample dimension\n\n" + _HELP)
        if self.channel_dim is None:
            raise RuntimeError("dataset does not have a channel dimension\n\n" + _HELP)
        if self.num_dims != len(shape):
            raise RuntimeError(
                f"dataset has {len(shape)} dimensions, but attribute axis_names has {self.num_dims} entries\n\n" + _HELP)

    @staticmethod
    def from_dataset_config(dataset_config: DatasetConfig) -> "DatasetMetaData":
        container = zarr_lite.open(dataset_config.container_path, "r")
        try:
            data = container[dataset_config.dataset_name]
        except KeyError:
            raise RuntimeError(
                f"Zarr container {dataset_config.container_path} does not contain "
                f'"{dataset_config.dataset_name}" dataset\n\n' + _HELP)
        try:
            axis_names = data.attrs["axis_names"]
        except KeyError:
            raise RuntimeError(
                f'"{dataset_config.dataset_name}" dataset in {dataset_config.container_path} does not contain '
                f'"axis_names" attribute\n\n' + _HELP)
        try:
            return DatasetMetaData(data.shape, axis_names)
        except RuntimeError as e:
            raise RuntimeError(
                f'"{dataset_config.dataset_name}" dataset in {dataset_config.container_path} has invalid meta-data'
            ) from e
