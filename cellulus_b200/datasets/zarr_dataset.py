"""Random-crop dataset over a zarr container (`cellulus/datasets/zarr_dataset.py`).

The reference builds a gunpowder graph (ZarrSource + RandomLocation + Normalize [+ ElasticAugment]); that
I/O library is out of scope (and not installed), so crops are cut directly from the array and the elastic
augmentation (rotation 0..pi/2, scale 0.9..1.1, control-point jitter) is applied by datasets/augment.py.  What IS on the hot
path -- the anchor / reference pair sampler (`:177-251`) -- is provided twice: `sample_coordinates()` is the
host sampler with the reference's exact RNG call order (what its DataLoader workers run), and
`sample_coordinates_device()` draws the same distribution on the GPU (`cb200_sample_pairs`), so the two
90 MB int64 lists of a config-#2 step never cross PCIe.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch
from torch.utils.data import IterableDataset

from cellulus_b200 import zarr_lite
from cellulus_b200.configs import DatasetConfig

from .augment import elastic_crop
from .meta_data import DatasetMetaData


def sample_offsets_within_radius(radius, number_offsets, num_spatial_dims):
    """`zarr_dataset.py:177-196` on numpy's global generator: `num_spatial_dims * number_offsets` candidates per
    axis, keep the open ball minus the origin, take the first `number_offsets`, draw again when short."""
    n = num_spatial_dims * number_offsets
    cols = [np.random.randint(-radius, radius + 1, size=n) for _ in range(num_spatial_dims)]
    offsets = np.stack(cols, axis=1)
    offsets = offsets[(offsets**2).sum(axis=1) < radius**2]
    offsets = offsets[np.absolute(offsets).sum(axis=1) > 0]
    if len(offsets) < number_offsets:
        return sample_offsets_within_radius(radius, number_offsets, num_spatial_dims)
    return offsets[:number_offsets]


def sample_coordinates(output_shape, kappa, num_anchors, num_references, num_spatial_dims):
    """`zarr_dataset.py:198-242`, same RNG call order (per-dim anchors, then per-dim offsets): int64 `(P, D)` anchor
    and reference lists of one sample, every anchor repeated `num_references` times (`np.repeat`, `:236`)."""
    cols = [np.random.randint(kappa, output_shape[d] - kappa + 1, size=num_anchors) for d in range(num_spatial_dims)]
    anchor_samples = np.repeat(np.stack(cols, axis=1), num_references, axis=0)
    reference_samples = anchor_samples + sample_offsets_within_radius(kappa, len(anchor_samples), num_spatial_dims)
    return anchor_samples, reference_samples


class ZarrDataset(IterableDataset):  # type: ignore
    def __init__(
        self,
        dataset_config: DatasetConfig,
        crop_size: Tuple[int, ...],
        elastic_deform: bool,
        control_point_spacing: int,
        control_point_jitter: float,
        density: float,
        kappa: float,
        normalization_factor: float,
        sample_pairs: bool = True,
        coordinate_dtype=np.int64,
    ):
        self.dataset_config = dataset_config
        self.crop_size = tuple(crop_size)
        self.elastic_deform = elastic_deform
        self.control_point_spacing = control_point_spacing
        self.control_point_jitter = control_point_jitter
        self.normalization_factor = normalization_factor
        self.sample_pairs = sample_pairs
        # host pair lists are int64 in the reference (numpy's default); a narrower type (np.int16 / np.int32) is
        # converted HERE, in the DataLoader workers, so that the training step ships a quarter / half of the bytes
        # over PCIe -- the loss kernels read int16 / int32 / int64 lists alike
        self.coordinate_dtype = np.dtype(coordinate_dtype)
        meta = DatasetMetaData.from_dataset_config(dataset_config)
        self.num_dims = meta.num_dims
        self.num_spatial_dims = meta.num_spatial_dims
        self.num_channels = meta.num_channels
        self.num_samples = meta.num_samples
        self.sample_dim, self.channel_dim, self.time_dim = meta.sample_dim, meta.channel_dim, meta.time_dim
        self.spatial_array = meta.spatial_array
        assert len(crop_size) == self.num_spatial_dims, (
            f'"crop_size" must have the same dimension as the spatial(temporal) dimensions of the '
            f'"{dataset_config.dataset_name}" dataset which is {self.num_spatial_dims}, but it is {crop_size}')
        self.density = density
        self.kappa = kappa
        self.output_shape = tuple(int(c - 16) for c in self.crop_size)  # hard-coded in the reference (:94)
        self.unbiased_shape = tuple(int(o - (2 * self.kappa)) for o in self.output_shape)

    def __iter__(self):
        return iter(self._yield_sample())

    def _normalize(self, data: np.ndarray, stored_dtype=None) -> np.ndarray:
        factor = self.normalization_factor
        if factor is None:  # gp.Normalize(factor=None): by the dtype the array is stored in
            factor = {np.dtype(np.uint8): 1.0 / 255, np.dtype(np.uint16): 1.0 / 65535}.get(
                np.dtype(stored_dtype if stored_dtype is not None else data.dtype), 1.0)
        return data.astype(np.float32) * np.float32(factor)

    def _crop(self, array, rng) -> np.ndarray:
        """One normalised crop `(C, *crop_size)`: a random location (gp.RandomLocation), elastically deformed when
        the config asks for it (`:122-131`; see datasets/augment.py)."""
        s = int(rng.integers(0, self.num_samples))
        if self.elastic_deform:
            data = elastic_crop(array, s, self.crop_size, self.spatial_array, self.control_point_spacing,
                                self.control_point_jitter, rng)
            return self._normalize(data, array.dtype)
        start = [int(rng.integers(0, n - c + 1)) for n, c in zip(self.spatial_array, self.crop_size)]
        key = (s, slice(None)) + tuple(slice(a, a + c) for a, c in zip(start, self.crop_size))
        return self._normalize(array[key])

    def _yield_sample(self):
        array = zarr_lite.open(self.dataset_config.container_path, "r")[self.dataset_config.dataset_name]
        rng = np.random.default_rng()
        while True:
            while True:  # reject all-zero crops (:138-151)
                crop = self._crop(array, rng)
                if np.max(crop) > 0.0:
                    break
            if self.sample_pairs:
                anchors, refs = self.sample_coordinates()
                if self.coordinate_dtype != anchors.dtype:
                    if max(self.output_shape) > np.iinfo(self.coordinate_dtype).max:
                        raise ValueError(f"coordinate_dtype {self.coordinate_dtype} cannot hold the output extent {self.output_shape}")
                    anchors, refs = anchors.astype(self.coordinate_dtype), refs.astype(self.coordinate_dtype)
                yield crop, anchors, refs
            else:
                yield crop

    # ---- the pair sampler (hot path)
    def sample_offsets_within_radius(self, radius, number_offsets):
        """`:177-196`: rejection-sample integer offsets in the open ball minus the origin."""
        return sample_offsets_within_radius(radius, number_offsets, self.num_spatial_dims)

    def sample_coordinates(self):
        """`:198-242`, same RNG call order: per-dim anchors, then per-dim offsets."""
        return sample_coordinates(self.output_shape, self.kappa, self.get_num_anchors(), self.get_num_references(),
                                  self.num_spatial_dims)

    def pair_stream(self):
        """Parameters of the device pair stream for this dataset's crops: what `criterion.fused_sampled` needs to
        draw the pairs inside the loss kernel (quirk Q4: column d is drawn from output_shape[d])."""
        return dict(kappa=self.kappa, num_anchors=self.get_num_anchors(), num_references=self.get_num_references(),
                    extent_xyz=tuple(self.output_shape[: self.num_spatial_dims]))

    def sample_coordinates_device(self, batch_size, device, seed, sequence=0, dtype=None):
        """The same distribution drawn on the GPU: (B, P, D) anchors and references.

        `dtype=None` picks the narrowest coordinate type that holds the output extent (int16 below 32768
        pixels per axis): the lists never leave the device, and the fused loss reads them at a quarter of the
        bytes of the reference's int64 lists.  Pass `torch.int64` for the reference's own format."""
        from cellulus_b200 import kernels as K

        if dtype is None:
            dtype = torch.int16 if max(self.output_shape[: self.num_spatial_dims]) < 2**15 else torch.int32

        return K.sample_pairs(batch_size, self.output_shape[: self.num_spatial_dims], self.kappa,
                              self.get_num_anchors(), self.get_num_references(), seed, sequence, dtype, device)

    def get_num_anchors(self):
        return int(self.density * self.unbiased_shape[0] * self.unbiased_shape[1])

    def get_num_references(self):
        return int(self.density * self.kappa**2 * np.pi)

    def get_num_samples(self):
        return self.get_num_anchors() * self.get_num_references()

    def get_num_channels(self):
        return self.num_channels

    def get_num_spatial_dims(self):
        return self.num_spatial_dims
