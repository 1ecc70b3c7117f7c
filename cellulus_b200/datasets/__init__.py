"""Dataset factory with the reference's call signature (`cellulus/datasets/__init__.py:8-27`)."""

from cellulus_b200.datasets.meta_data import DatasetMetaData  # noqa: F401
from cellulus_b200.datasets.staging import PairListStager  # noqa: F401
from cellulus_b200.datasets.zarr_dataset import ZarrDataset


def get_dataset(dataset_config, crop_size, elastic_deform, control_point_spacing, control_point_jitter, density,
                kappa, normalization_factor) -> ZarrDataset:
    """Random-crop dataset over a zarr container; every argument is handed to `ZarrDataset` by name."""
    arguments = dict(locals())
    return ZarrDataset(**arguments)
