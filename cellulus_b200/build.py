"""Build `libcellulus_b200.so` (the C-ABI CUDA library) in-tree with nvcc.

    python -m cellulus_b200.build [--force]

Every `.cu` under `cellulus_b200/csrc/` is compiled for sm_100a only
(`-gencode arch=compute_100a,code=sm_100a -lineinfo`) and linked into ONE shared
object next to the package, so the built library travels with the repository
snapshot.  No torch headers are involved: the ABI is plain C (include/cellulus_b200.h).
"""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "libcellulus_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libcellulus_b200.so")
    return nvcc


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    with open(os.path.join(os.path.dirname(PKG), "include", "cellulus_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    stamp = LIB + ".digest"
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == _digest()


def build(force: bool = False, verbose: bool = True, extra_flags=(), out: str | None = None) -> str:
    """`extra_flags` / `out` build a tuning variant (e.g. -DCB200_LOSS_UNROLL=4) next to the real library."""
    variant = bool(extra_flags) or out is not None
    if not variant and not force and is_fresh():
        return LIB
    build_dir = BUILD if not variant else os.path.join(BUILD, "variant_" + hashlib.sha1(" ".join(extra_flags).encode()).hexdigest()[:8])
    os.makedirs(build_dir, exist_ok=True)
    nvcc = _nvcc()
    lib_out = out or LIB

    def compile_one(src):
        obj = os.path.join(build_dir, src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", lib_out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if not variant:
        with open(LIB + ".digest", "w") as fh:
            fh.write(_digest())
    if verbose:
        print(f"built {lib_out} from {len(objs)} translation units", file=sys.stderr)
    return lib_out


if __name__ == "__main__":
    build(force="--force" in sys.argv)
