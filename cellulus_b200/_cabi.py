"""ctypes binding of `libcellulus_b200.so` (C ABI declared in include/cellulus_b200.h).

The product path has NO CPU fallback: if the library is missing the import of
any op raises, loudly, with the build command.
"""

from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CELLULUS_B200_LIB") or os.path.join(PKG, "libcellulus_b200.so")  # override: A/B builds

OK, EINVAL, EUNSUPPORTED, ENOFIT, ENOCENTRE, ENOCONVERGE, ENOSPACE = 0, -1, -2, -3, -4, -5, -6
F32, BF16, F64, I64, I32, I16, U8, U16 = range(8)
LAYOUT_PLANAR, LAYOUT_CHANNELS_LAST = 0, 1


class Grid(C.Structure):
    """`cb200_grid` (include/cellulus_b200.h)."""

    _fields_ = [
        ("origin", C.c_double * 3),
        ("cell", C.c_double),
        ("inv_cell", C.c_double),
        ("dims", C.c_int32 * 3),
        ("num_dims", C.c_int32),
        ("n_cells", C.c_int64),
    ]


class DetectInfo(C.Structure):
    """`cb200_detect_info` (include/cellulus_b200.h)."""

    _fields_ = [
        ("n_foreground", C.c_int64),
        ("n_fit", C.c_int64),
        ("n_seeds", C.c_int64),
        ("n_centres", C.c_int32),
        ("suppress_calls", C.c_int32),
        ("grid", Grid),
        ("distance_tests", C.c_int64),
        ("climb_steps", C.c_int64),
        ("workspace_needed", C.c_int64),
        ("n_distinct_modes", C.c_int64),
    ]


class CellulusB200Error(RuntimeError):
    pass


_p, _i, _i64, _f, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol the header declares
PROTOTYPES = {
    "cb200_version": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "cb200_error_string": (C.c_char_p, [_i]),
    "cb200_fma_peak": (_i, [_i, _i, _i, _p, _pi64, _p]),
    "cb200_oce_loss_workspace_bytes": (_i64, []),
    "cb200_oce_loss_staging_bytes": (_i64, [_i, _i, _i, _i, _pi64]),
    "cb200_oce_loss_fwd_bwd": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _pi64, _i64, _f, _f, _p, _p, _p, _p]),
    "cb200_oce_loss_fwd_bwd_staged": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _pi64, _i64, _f, _f, _p, _p, _p, _p, _i64, _p]),
    "cb200_scale_inplace": (_i, [_p, _i64, _p, _p]),
    "cb200_gather_add_coords": (_i, [_p, _i, _p, _i, _i, _i, _pi64, _i64, _p, _p]),
    "cb200_scatter_add_coords": (_i, [_p, _p, _i, _i, _i, _pi64, _i64, _p, _p]),
    "cb200_oce_pair_loss": (_i, [_p, _p, _i64, _i, _f, _f, _p, _p, _p, _p]),
    "cb200_sample_pairs": (_i, [_p, _p, _i, _i, _i, _pi64, _d, _i64, _i, _u64, _u64, _p]),
    "cb200_oce_loss_sampled": (_i, [_p, _i, _i, _i, _i, _pi64, _pi64, _d, _i64, _i, _u64, _u64, _f, _f, _p, _p, _p, _p, _p,
                                    _i, _p]),
    "cb200_oce_loss_sampled_staged": (_i, [_p, _i, _i, _i, _i, _pi64, _pi64, _d, _i64, _i, _u64, _u64, _f, _f, _p, _p, _p,
                                           _p, _p, _i, _p, _i64, _p]),
    "cb200_tta_aggregate": (_i, [_p, _i, _i, _i64, _p, _p]),
    "cb200_tta_accumulate": (_i, [_p, _p, _i, _i, _i64, _p]),
    "cb200_tta_finalize": (_i, [_p, _i, _i, _i64, _p, _p]),
    "cb200_salt_pepper": (_i, [_p, _i64, _f, _f, _u64, _u64, _p, _p]),
    "cb200_salt_pepper_device_seed": (_i, [_p, _i64, _f, _f, _p, _u64, _p, _p]),
    "cb200_centre_workspace_bytes": (_i64, []),
    "cb200_centre_embeddings": (_i, [_p, _i, _i, _i64, _d, _p, _p, _p, _p]),
    "cb200_channel_norm": (_i, [_p, _i, _i, _i64, _p, _p]),
    "cb200_gaussian_blur": (_i, [_p, _p, _p, _i, _pi64, C.POINTER(_d), _i, _i, _p]),
    "cb200_peaks_workspace_bytes": (_i64, [_i64]),
    "cb200_local_peaks": (_i, [_p, _i, _pi64, _d, _p, _p, _i64, _p, _p, _p]),
    "cb200_reduce_workspace_bytes": (_i64, []),
    "cb200_minmax": (_i, [_p, _i, _i64, _p, _p, _p]),
    "cb200_histogram": (_i, [_p, _i, _i64, _p, _i, _p, _p]),
    "cb200_compact_workspace_bytes": (_i64, [_i64]),
    "cb200_fg_compact": (_i, [_p, _i, _i, _pi64, _d, _p, _p, _i64, _p, _p, _i, _p, _p]),
    "cb200_select_points": (_i, [_p, _i64, _i64, _i, _p, _p, _i64, _p, _p, _p]),
    "cb200_bernoulli_flags": (_i, [_p, _i64, _d, _u64, _p]),
    "cb200_ms_brute_partial_bytes": (_i64, [_i64, _i64, _i]),
    "cb200_ms_brute_accumulate": (_i, [_p, _i64, _i64, _i, _p, _i64, _p, _i64, _d, _p, _p]),
    "cb200_ms_update": (_i, [_p, _i64, _i, _p, _p, _p, _i64, _i64, _p, _d, _i, _p, _p, _p]),
    "cb200_grid_plan": (_i, [C.POINTER(_d), C.POINTER(_d), _i, _d, _i64, C.POINTER(Grid)]),
    "cb200_grid_build_workspace_bytes": (_i64, [_i64, _i64]),
    "cb200_grid_build": (_i, [_p, _i64, _i64, C.POINTER(Grid), _p, _i64, _p, _p, _p, _i64, _p]),
    "cb200_ms_grid_modes": (_i, [_p, _i64, _i64, C.POINTER(Grid), _p, _p, _i64, _i64, _d, _i, _p, _p, _p, _p]),
    "cb200_ms_distinct_workspace_bytes": (_i64, [_i64]),
    "cb200_ms_grid_modes_distinct": (_i, [_p, _i64, _i64, C.POINTER(Grid), _p, _p, _i64, _i64, _d, _i, _i, _p, _p, _p, _p,
                                          _i64, _p]),
    "cb200_bin_seeds_workspace_bytes": (_i64, [_i64]),
    "cb200_bin_seeds": (_i, [_p, _i64, _i64, _i, _d, _p, _i64, _p, _p, _p, _i64, _p]),
    "cb200_unique_modes_workspace_bytes": (_i64, [_i64]),
    "cb200_unique_modes": (_i, [_p, _i64, _i, _p, _i64, _p, _i64, _p, _p, _p, _i64, _p]),
    "cb200_nms_workspace_bytes": (_i64, [_i64, C.POINTER(Grid), _d]),
    "cb200_nms_suppress": (_i, [_p, _i64, _i, _p, _i64, _d, C.POINTER(Grid), _i, _i, _p, _p, _i64, _p]),
    "cb200_nms_emit": (_i, [_p, _i64, _i, _p, _i64, _d, C.POINTER(Grid), _i, _p, _i64, _p, _i64, _p]),
    "cb200_assign_workspace_bytes": (_i64, [_i64, _i, _i64]),
    "cb200_assign_labels": (_i, [_p, _i64, _i64, _i, _p, _i64, _i, C.POINTER(Grid), _p, _p, _i, _p, _p]),
    "cb200_greedy_workspace_bytes": (_i64, [_i64]),
    "cb200_greedy_prepare": (_i, [_p, _i, _i, _pi64, _p, _i, _d, _d, _p, _i64, _p, _p, _p, _p, _p]),
    "cb200_greedy_cluster": (_i, [_p, _i64, _p, _i64, _i, _i, _d, _i, _d, C.c_longlong, _p, _p, _p, _p]),
    "cb200_scatter_i16": (_i, [_p, _p, _i64, _p, _p]),
    "cb200_cc_workspace_bytes": (_i64, [_i64]),
    "cb200_label_components": (_i, [_p, _i, _pi64, _p, _p, _p, _p]),
    "cb200_size_filter": (_i, [_p, _i, _pi64, _i64, _p, _p, _p, _p]),
    "cb200_edt_workspace_bytes": (_i64, [_i64]),
    "cb200_edt_within": (_i, [_p, _i, _pi64, _d, _p, _p, _p]),
    "cb200_grow_shrink": (_i, [_p, _i, _pi64, _d, _d, _p, _p]),
    "cb200_detect_volume_workspace_bytes": (_i64, [_i, _pi64, _i64, _d]),
    "cb200_detect_volume": (_i, [_p, _i, _i, _pi64, _d, _d, _d, _u64, _i, _p, _i, _p, _i, _p, _i64, _p, _i64, _i64,
                                 C.POINTER(DetectInfo), _p]),
    "cb200_label_presence": (_i, [_p, _i, _i64, _i, _p, _p]),
    "cb200_contingency": (_i, [_p, _p, _i, _i64, _p, _p, _i, _i, _i, _p, _p]),
    "cb200_label_stats_workspace_bytes": (_i64, [_i]),
    "cb200_label_stats": (_i, [_p, _p, _i, _i, _pi64, _i, _p, _p, _p, _p, _p]),
    "cb200_label_histogram": (_i, [_p, _p, _i, _i64, _i, _p, _p, _p, _i, _p, _p]),
    "cb200_label_otsu_workspace_bytes": (_i64, [_i64]),
    "cb200_label_otsu": (_i, [_p, _p, _p, _p, _p, _i, _i, _i64, _p, _p, _p]),
    "cb200_nucleus_fill_workspace_bytes": (_i64, [_i64, _i]),
    "cb200_nucleus_fill": (_i, [_p, _p, _i, _i, _pi64, _i, _p, _p, _p, _p, _i64, _p, _p, _p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CellulusB200Error(
            f"{LIB_PATH} is missing: the CUDA extension has not been built and there is no CPU fallback. "
            "Run `python -m cellulus_b200.build` (needs nvcc; cross-compiles sm_100a without a GPU)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code == OK:
        return
    msg = load().cb200_error_string(code)
    raise CellulusB200Error(f"{what} failed: {msg.decode() if msg else code} (code {code})")


def spatial_array(shape):
    return (C.c_int64 * len(shape))(*[int(s) for s in shape])
