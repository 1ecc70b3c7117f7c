"""Minimal zarr-v2 directory store -- only what the cellulus dataset layout needs.

The reference reads / writes zarr containers (`datasets/meta_data.py:50-76`, `predict.py:103-112,137-142`,
`detect.py:20-80`, `segment.py:20-38`): N-d arrays `(s, c, [z,] y, x)` with an `axis_names` attribute.
`zarr` is not installed in this image, so this module implements the on-disk format directly: groups are
directories with `.zgroup`, arrays are directories with `.zarray` (zarr_format 2, C order, "." chunk-key
separator), attributes live in `.zattrs`, one file per chunk.  Chunks are stored raw (`compressor: null`) or
zlib/gzip-compressed; containers written with blosc (the zarr default) need the real library:
`open()` hands over to `zarr.open` whenever `import zarr` succeeds.
"""

from __future__ import annotations

import builtins
import itertools
import json
import os
import zlib

import numpy as np

_fopen = builtins.open  # this module defines its own `open` (mirroring zarr.open)


class Attributes(dict):
    def __init__(self, path):
        super().__init__()
        self._path = path
        if os.path.exists(path):
            with _fopen(path) as fh:
                super().update(json.load(fh))

    def _flush(self):
        with _fopen(self._path, "w") as fh:
            json.dump(dict(self), fh, indent=2, default=_jsonable)

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self._flush()

    def update(self, *args, **kwargs):
        super().update(*args, **kwargs)
        self._flush()


def _jsonable(obj):
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if isinstance(obj, (tuple, np.ndarray)):
        return list(obj)
    raise TypeError(f"not JSON serialisable: {type(obj)}")


def _default_chunks(shape, dtype):
    """~4 MB chunks, whole trailing (spatial) axes first."""
    chunks = list(shape)
    limit = max(1, (4 << 20) // np.dtype(dtype).itemsize)
    for axis in range(len(shape)):
        if int(np.prod(chunks)) <= limit:
            break
        chunks[axis] = max(1, int(chunks[axis] // max(1, int(np.prod(chunks)) // limit)))
    return tuple(max(1, c) for c in chunks)


class Array:
    def __init__(self, path):
        self.path = path
        with _fopen(os.path.join(path, ".zarray")) as fh:
            meta = json.load(fh)
        if meta.get("zarr_format") != 2:
            raise ValueError(f"{path}: only zarr_format 2 is supported")
        if meta.get("order", "C") != "C" or meta.get("filters"):
            raise ValueError(f"{path}: only C order without filters is supported")
        comp = meta.get("compressor")
        if comp is not None and comp.get("id") not in ("zlib", "gzip"):
            raise ValueError(
                f"{path}: compressor {comp.get('id')!r} needs the real `zarr`/`numcodecs` packages "
                "(this minimal store reads raw, zlib and gzip chunks)")
        self._comp = comp
        self.shape = tuple(meta["shape"])
        self.chunks = tuple(meta["chunks"])
        self.dtype = np.dtype(meta["dtype"])
        self.fill_value = meta.get("fill_value", 0) or 0
        self._sep = meta.get("dimension_separator", ".")
        self.attrs = Attributes(os.path.join(path, ".zattrs"))

    ndim = property(lambda self: len(self.shape))

    def __len__(self):
        return self.shape[0]

    # ---- chunk I/O
    def _chunk_path(self, idx):
        return os.path.join(self.path, self._sep.join(str(i) for i in idx) if idx else "0")

    def _read_chunk(self, idx):
        p = self._chunk_path(idx)
        if not os.path.exists(p):
            return np.full(self.chunks, self.fill_value, dtype=self.dtype)
        with _fopen(p, "rb") as fh:
            raw = fh.read()
        if self._comp is not None:
            raw = zlib.decompress(raw, 15 + 32)
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks).copy()

    def _write_chunk(self, idx, data):
        raw = np.ascontiguousarray(data, dtype=self.dtype).tobytes()
        if self._comp is not None:
            level = int(self._comp.get("level", 1))
            if self._comp.get("id") == "gzip":  # numcodecs GZip.decode needs a gzip container, not a bare zlib stream
                co = zlib.compressobj(level, zlib.DEFLATED, 31)
                raw = co.compress(raw) + co.flush()
            else:
                raw = zlib.compress(raw, level)
        with _fopen(self._chunk_path(idx), "wb") as fh:
            fh.write(raw)

    # ---- selection handling: ints, slices, Ellipsis (step 1) -- what the reference uses
    def _normalise(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        if any(k is Ellipsis for k in key):
            i = key.index(Ellipsis)
            key = key[:i] + (slice(None),) * (self.ndim - len(key) + 1) + key[i + 1:]
        key = key + (slice(None),) * (self.ndim - len(key))
        sel, squeeze = [], []
        for axis, (k, n) in enumerate(zip(key, self.shape)):
            if isinstance(k, (int, np.integer)):
                k = int(k) + (n if k < 0 else 0)
                if not 0 <= k < n:
                    raise IndexError(f"index {k} out of bounds for axis {axis} with size {n}")
                sel.append((k, k + 1))
                squeeze.append(axis)
            elif isinstance(k, slice):
                start, stop, step = k.indices(n)
                if step != 1:
                    raise IndexError("only unit-step slices are supported")
                sel.append((start, max(start, stop)))
            else:
                raise IndexError(f"unsupported index {k!r}")
        return sel, tuple(squeeze)

    def _chunk_ranges(self, sel):
        per_axis = []
        for (lo, hi), c in zip(sel, self.chunks):
            per_axis.append(range(lo // c, (max(hi, lo + 1) - 1) // c + 1) if hi > lo else range(0))
        return itertools.product(*per_axis)

    def __getitem__(self, key):
        sel, squeeze = self._normalise(key)
        out = np.empty([hi - lo for lo, hi in sel], dtype=self.dtype)
        for idx in self._chunk_ranges(sel):
            chunk = self._read_chunk(idx)
            src, dst = [], []
            for i, (lo, hi), c in zip(idx, sel, self.chunks):
                a, b = max(lo, i * c), min(hi, (i + 1) * c)
                src.append(slice(a - i * c, b - i * c))
                dst.append(slice(a - lo, b - lo))
            out[tuple(dst)] = chunk[tuple(src)]
        return out.squeeze(axis=squeeze) if squeeze else out

    def __setitem__(self, key, value):
        sel, squeeze = self._normalise(key)
        shape = [hi - lo for lo, hi in sel]
        value = np.asarray(value)
        if squeeze and value.ndim == len(shape) - len(squeeze):
            value = np.expand_dims(value, squeeze)
        value = np.broadcast_to(value.astype(self.dtype, copy=False), shape)
        for idx in self._chunk_ranges(sel):
            src, dst, full = [], [], True
            for i, (lo, hi), c, n in zip(idx, sel, self.chunks, self.shape):
                a, b = max(lo, i * c), min(hi, (i + 1) * c)
                dst.append(slice(a - i * c, b - i * c))
                src.append(slice(a - lo, b - lo))
                full = full and (b - a == c or (a == i * c and b == n))
            if full and all(d.start == 0 for d in dst) and all(
                    s.stop - s.start == c for s, c in zip(src, self.chunks)):
                chunk = value[tuple(src)]
            else:
                chunk = self._read_chunk(idx)
                chunk[tuple(dst)] = value[tuple(src)]
            self._write_chunk(idx, chunk)


class Group:
    def __init__(self, path, mode="a"):
        self.path = str(path)
        self.mode = mode
        if not os.path.isdir(self.path):
            if mode == "r":
                raise FileNotFoundError(self.path)
            os.makedirs(self.path, exist_ok=True)
        marker = os.path.join(self.path, ".zgroup")
        if not os.path.exists(marker) and mode != "r":
            with _fopen(marker, "w") as fh:
                json.dump({"zarr_format": 2}, fh)
        self.attrs = Attributes(os.path.join(self.path, ".zattrs"))

    def __contains__(self, name):
        return os.path.exists(os.path.join(self.path, name, ".zarray")) or os.path.exists(
            os.path.join(self.path, name, ".zgroup"))

    def __getitem__(self, name):
        p = os.path.join(self.path, name)
        if os.path.exists(os.path.join(p, ".zarray")):
            return Array(p)
        if os.path.exists(os.path.join(p, ".zgroup")):
            return Group(p, self.mode)
        raise KeyError(name)

    def create_dataset(self, name, shape, dtype=float, chunks=None, overwrite=False, compressor=None, **_):
        p = os.path.join(self.path, name)
        if os.path.exists(os.path.join(p, ".zarray")) and not overwrite:
            raise ValueError(f"path {name!r} contains an array")  # zarr v2 ContainsArrayError semantics
        parent = self
        parts = name.strip("/").split("/")
        for part in parts[:-1]:  # intermediate groups, e.g. snapshots "1000/raw"
            parent = Group(os.path.join(parent.path, part), "a")
        os.makedirs(p, exist_ok=True)
        for f in os.listdir(p):
            if not f.startswith(".z") and overwrite:
                os.remove(os.path.join(p, f))
        dtype = np.dtype(dtype)
        shape = tuple(int(s) for s in shape)
        meta = {
            "zarr_format": 2, "shape": list(shape),
            "chunks": list(chunks if chunks is not None else _default_chunks(shape, dtype)),
            "dtype": dtype.str, "compressor": compressor, "fill_value": 0, "order": "C", "filters": None,
        }
        with _fopen(os.path.join(p, ".zarray"), "w") as fh:
            json.dump(meta, fh, indent=2)
        return Array(p)

    def __setitem__(self, name, value):
        value = np.asarray(value)
        arr = self.create_dataset(name, value.shape, value.dtype, overwrite=True)
        arr[...] = value


def open(path, mode="a"):  # noqa: A001 - mirrors zarr.open
    """`zarr.open(path, mode)`: the real library if it is installed, this minimal store otherwise."""
    try:
        import zarr  # type: ignore

        return zarr.open(str(path), mode=mode)
    except ImportError:
        return Group(path, mode)
