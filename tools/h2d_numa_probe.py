"""Probe (torchrun, one rank per GPU): aggregate pinned-host -> device bandwidth of N ranks copying 195.6 MB each
per step (the e2e loss step of bench.py), with the pinned buffers placed (a) wherever the default CPU affinity
puts them and (b) spread over the host's NUMA nodes by rank (first touch after sched_setaffinity).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_numa_probe.py
"""
import glob
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def node_cpus():
    nodes = {}
    for path in sorted(glob.glob("/sys/devices/system/node/node[0-9]*/cpulist")):
        node = int(path.split("node")[-1].split("/")[0])
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        nodes[node] = cpus
    return nodes


base_affinity = os.sched_getaffinity(0)
nodes = node_cpus()
if rank == 0:
    print("numa nodes:", {k: len(v) for k, v in nodes.items()}, "affinity:", len(base_affinity), flush=True)
    try:
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout, flush=True)
    except Exception as e:  # noqa: BLE001
        print("nvidia-smi topo failed:", e)

NBYTES = 195_550_976
for mode in ("default", "spread", "node_last"):
    os.sched_setaffinity(0, base_affinity)
    if mode != "default" and len(nodes) > 1:
        keys = sorted(nodes)
        node = keys[rank % len(keys)] if mode == "spread" else keys[-1]
        allowed = nodes[node] & base_affinity
        if allowed:
            os.sched_setaffinity(0, allowed)
    host = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
    host.fill_(1)  # touch
    dst = torch.empty(NBYTES, dtype=torch.uint8, device=dev)
    for _ in range(3):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    sec = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{mode:10s} per-rank {NBYTES * 20 / sec.item() / 1e9:6.1f} GB/s  aggregate {world * NBYTES * 20 / sec.item() / 1e9:7.1f} GB/s "
              f"({sec.item() / 20 * 1e3:.2f} ms per 195.6 MB)", flush=True)
    del host, dst
if world > 1:
    dist.destroy_process_group()
