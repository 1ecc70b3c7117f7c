"""CPU: the C-ABI library builds, loads and exports every symbol include/cellulus_b200.h declares
(no compute calls: there is no GPU here)."""

import ctypes
import os
import re

import pytest

from cellulus_b200 import _cabi, build

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(REPO, "include", "cellulus_b200.h")).read()
    return sorted(set(re.findall(r"CB200_API\s+[\w\s\*]+?\b(cb200_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    build.build(verbose=False)  # no-op when the in-tree .so is fresh
    return _cabi.load()


def test_header_and_binding_agree():
    declared = _declared()
    assert len(declared) >= 30
    assert declared == sorted(_cabi.PROTOTYPES)


def test_every_symbol_is_exported(lib):
    for name in _declared():
        assert hasattr(lib, name), name
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared():
        getattr(raw, name)


def test_host_only_entry_points(lib):
    major, minor, arch = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.cb200_version(major, minor, arch) == 0
    assert arch.value == 100
    assert lib.cb200_error_string(-1) == b"invalid argument"
    assert lib.cb200_oce_loss_workspace_bytes() >= 24
    assert lib.cb200_compact_workspace_bytes(1 << 20) > 0
    assert lib.cb200_cc_workspace_bytes(1 << 20) >= 8 << 20
    g = _cabi.Grid()
    lo = (ctypes.c_double * 3)(0.0, 0.0, 0.0)
    hi = (ctypes.c_double * 3)(100.0, 50.0, 25.0)
    assert lib.cb200_grid_plan(lo, hi, 3, 5.0, 1 << 26, ctypes.byref(g)) == 0
    assert g.cell >= 5.0 and list(g.dims) == [20, 10, 5] and g.n_cells == 1000
    # max_cells forces coarser (still valid) cells
    assert lib.cb200_grid_plan(lo, hi, 3, 5.0, 100, ctypes.byref(g)) == 0
    assert g.n_cells <= 100 and g.cell > 5.0
    assert lib.cb200_grid_plan(lo, hi, 4, 5.0, 0, ctypes.byref(g)) == _cabi.EINVAL


def test_sass_is_sm100a_with_bulk_copy():
    """The library holds sm_100a code only, and the n-body kernel really uses the bulk async copy engine."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    build.build(verbose=False)
    elf = subprocess.run([cuobjdump, "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", elf))
    assert archs == {"100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "ms_brute_accumulate_kernel", _cabi.LIB_PATH],
                          capture_output=True, text=True).stdout
    if "UBLKCP" not in sass:  # -fun needs the mangled name on some versions: fall back to a full dump
        sass = subprocess.run([cuobjdump, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass


def test_cpu_tensors_are_refused(lib):
    import torch

    from cellulus_b200 import kernels as K
    from cellulus_b200.criterions import get_loss

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K.tta_aggregate(torch.zeros(4, 2, 8, 8))
    crit = get_loss(temperature=10.0, regularizer_weight=1e-5, density=0.1, num_spatial_dims=2, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libcellulus_b200.so")
    with pytest.raises(_cabi.CellulusB200Error, match="no CPU fallback"):
        _cabi.load()
