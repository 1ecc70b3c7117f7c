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


def test_arguments_are_validated_before_any_device_work(lib):
    """Bad arguments come back as status codes (never exceptions, never a launch): checked here without a GPU."""
    sp2 = (ctypes.c_int64 * 2)(16, 16)
    sp4 = (ctypes.c_int64 * 4)(2, 2, 2, 2)
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    info = _cabi.DetectInfo()
    assert lib.cb200_edt_within(None, 2, sp2, 3.0, p, p, None) == _cabi.EINVAL
    assert lib.cb200_edt_within(p, 4, sp4, 3.0, p, p, None) == _cabi.EUNSUPPORTED
    assert lib.cb200_grow_shrink(None, 2, sp2, 3.0, 6.0, p, None) == _cabi.EINVAL
    assert lib.cb200_label_stats(p, p, _cabi.F32, 1, sp2, 5, p, p, p, p, None) == _cabi.EUNSUPPORTED
    assert lib.cb200_label_otsu(p, p, p, None, None, _cabi.F64, 3, 10, p, p, None) == _cabi.EINVAL
    assert lib.cb200_contingency(p, p, _cabi.U16, 10, p, p, -1, 2, 2, p, None) == _cabi.EINVAL
    assert lib.cb200_detect_volume(p, _cabi.F32, 2, sp2, 0.5, 0.0, 1.0, 0, 300, p, _cabi.I32, None, 0, None, 0, p, 64, 0,
                                   ctypes.byref(info), None) == _cabi.EINVAL  # bandwidth must be positive
    assert lib.cb200_detect_volume(p, _cabi.F32, 2, sp2, 0.5, 3.0, 1.0, 0, 300, p, _cabi.I32, None, 0, None, 0, None, 0, 0,
                                   ctypes.byref(info), None) == _cabi.EINVAL  # the scratch is the caller's
    assert lib.cb200_detect_volume_workspace_bytes(2, sp2, 0, 1.0) > lib.cb200_detect_volume_workspace_bytes(2, sp2, 16, 0.1) > 0
    assert lib.cb200_detect_volume(p, _cabi.F32, 4, sp4, 0.5, 3.0, 1.0, 0, 300, p, _cabi.I32, None, 0, None, 0, p, 64, 0,
                                   ctypes.byref(info), None) == _cabi.EUNSUPPORTED
    assert lib.cb200_sample_pairs(p, p, _cabi.F32, 1, 2, sp2, 3.0, 4, 4, 0, 0, None) == _cabi.EUNSUPPORTED
    assert lib.cb200_error_string(_cabi.ENOFIT) == b"the fit subset is empty"
    assert lib.cb200_nucleus_fill_workspace_bytes(1000, 7) >= 4 * 1007
    assert lib.cb200_edt_workspace_bytes(1 << 20) >= 9 << 20


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
