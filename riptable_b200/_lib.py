"""ctypes binding of libriptable_b200.so (the C ABI declared in include/riptable_b200.h).

The product path has no CPU fallback: if the CUDA library is missing this module raises at import.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RTB_LIB: measurement aid (profiles/r02/ab: an older build of this library timed beside the current one)
LIB_PATH = os.environ.get("RTB_LIB") or os.path.join(HERE, "libriptable_b200.so")

RTB_OK, RTB_ERR_ARG, RTB_ERR_WORKSPACE, RTB_ERR_CUDA, RTB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4


class rtb_column(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("itemsize", C.c_int32)]


class rtb_gb_state(C.Structure):
    _fields_ = [("unique", C.c_int64), ("prev_new", C.c_int64), ("prev_rows", C.c_int64), ("max_unique", C.c_int64),
                ("cap", C.c_uint32), ("rounds", C.c_int32)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"riptable_b200: {LIB_PATH} is missing. Build it with `python -m riptable_b200.build` "
        "(nvcc, sm_100a). There is no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_vp, _i64, _i32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_size_t
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol of include/riptable_b200.h
PROTOTYPES = {
    "rtb_version": (C.c_int, []),
    "rtb_last_error": (C.c_char_p, []),
    "rtb_launch_count": (C.c_uint64, []),
    "rtb_profile_enable": (None, [_i32]),
    "rtb_profile_reset": (None, []),
    "rtb_profile_get": (C.c_int, [_i32, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "rtb_groupby_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "rtb_groupby_workspace_bytes_rows": (_sz, [_i64, _i64, _i64, _i32]),
    "rtb_multikey_groupby32": (
        C.c_int,
        [C.POINTER(rtb_column), _i32, _i64, _i64, _vp, _pi64, _i64, _vp, _vp, _i64, _pi64, _pi64, _pi64, _vp, _sz, _vp],
    ),
    "rtb_multikey_groupby32_chunk": (
        C.c_int, [C.POINTER(rtb_column), _i64, _vp, _vp, _vp, _i64, _i64, _vp, _i32, _pi64, _vp, _sz, _vp]),
    "rtb_groupby_sum_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "rtb_multikey_groupby_sum32": (
        C.c_int,
        [C.POINTER(rtb_column), _i32, _i64, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _i64, _vp, _pi64, _pi64, _vp, _sz, _vp],
    ),
    "rtb_groupby_sum_chunk_workspace_bytes": (_sz, [_i64, _i64]),
    "rtb_multikey_groupby_sum32_chunk": (
        C.c_int,
        [C.POINTER(rtb_column), _i64, _vp, _vp, _i32, _i32, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i32, _pi64, _vp, _sz, _vp],
    ),
    "rtb_ismember_categorical_fixup": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp, _vp]),
    "rtb_first_row_per_code": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _vp]),
    "rtb_multikey_groupby32_adopt_seed": (C.c_int, [C.POINTER(rtb_column), _i64, _vp, _i64, _vp, _vp, _vp, _pi64, _vp, _sz, _vp]),
    "rtb_positive_mask": (C.c_int, [_vp, _i32, _i64, _vp, _vp]),
    "rtb_ismember_finish": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp]),
    "rtb_combine_filter": (C.c_int, [_vp, _i32, _vp, _i64, _vp, _vp]),
    "rtb_combine_accum2": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _i64, _i64, _vp, _i32, _vp]),
    "rtb_reduce_out_dtype": (C.c_int, [_i32, _i32]),
    "rtb_groupby_reduce_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "rtb_groupby_reduce": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i64, _i32, _i64, _i64, _vp, _vp, _sz, _vp]),
    "rtb_bincount": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _vp]),
    "rtb_pack_workspace_bytes": (_sz, [_i64, _i64]),
    "rtb_groupby_pack": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rtb_groupby_pick": (C.c_int, [_vp, _i32, C.c_int32, _vp, _vp, _vp, _i64, _i32, _i64, _i64, _i64, _vp, _vp]),
    "rtb_cumsum_out_dtype": (C.c_int, [_i32]),
    "rtb_reindex_groups": (C.c_int, [_vp, _i64, _vp, _i64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _i64, _vp, _vp]),
    "rtb_quantile_workspace_bytes": (_sz, [_i64, _i64]),
    "rtb_groupby_quantile": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _i64, _i64, C.c_double, _i32, _i64, _i64, _vp, _vp, _sz, _vp]),
    "rtb_ema_out_dtype": (C.c_int, [_i32, _i32]),
    "rtb_ema_workspace_bytes": (_sz, [_i64, _i64]),
    "rtb_groupby_ema": (C.c_int, [_i32, _vp, _i32, _vp, _i32, _i64, _i64, _vp, _i32, C.c_double, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rtb_ismember_workspace_bytes": (_sz, [_i64]),
    "rtb_ismember32": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "rtb_groupby_mode": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _sz, _vp]),
    "rtb_groupby_rolling_quantile": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, C.c_double, _vp, _vp]),
    "rtb_rolling_out_dtype": (C.c_int, [_i32, _i32]),
    "rtb_groupby_rolling": (C.c_int, [_i32, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "rtb_cumop_out_dtype": (C.c_int, [_i32, _i32]),
    "rtb_groupby_cumop": (C.c_int, [_i32, _vp, _i32, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rtb_cumsum_workspace_bytes": (_sz, [_i64, _i64]),
    "rtb_cumop_workspace_bytes": (_sz, [_i32, _i64, _i64, _i32]),
    "rtb_groupby_cumsum": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "rtb_make_ifirst": (C.c_int, [_vp, _i32, _i64, _i64, _vp, _i32, _vp, _vp]),
    "rtb_make_inext_workspace_bytes": (_sz, [_i64, _i64]),
    "rtb_make_inext": (C.c_int, [_vp, _i32, _i64, _i64, _i32, _vp, _vp, _sz, _vp]),
    "rtb_lexsort_workspace_bytes": (_sz, [_i64, _i64, _i32, C.POINTER(C.c_int32)]),
    "rtb_lexsort32": (C.c_int, [C.POINTER(rtb_column), _i32, _i64, _vp, _i64, C.POINTER(C.c_int8), _vp, _vp, _sz, _vp]),
    "rtb_lexsort64_workspace_bytes": (_sz, [_i64, _i64, _i32, C.POINTER(C.c_int32)]),
    "rtb_lexsort64": (C.c_int, [C.POINTER(rtb_column), _i32, _i64, _vp, _i64, C.POINTER(C.c_int8), _vp, _vp, _sz, _vp]),
    "rtb_lex_partition": (C.c_int, [C.POINTER(rtb_column), _i32, _i64, C.POINTER(C.c_int8), C.POINTER(rtb_column), _i32, _vp,
                                    _i64, _vp, _vp]),
    "rtb_group_from_lexsort_workspace_bytes": (_sz, [_i64, _i64]),
    "rtb_group_from_lexsort": (
        C.c_int,
        [_vp, _i64, C.POINTER(rtb_column), _i32, _i64, _i32, _vp, _vp, _vp, _pi64, _vp, _sz, _vp],
    ),
    "rtb_partition_ids": (C.c_int, [_i64, _vp, _i64, _vp, _vp]),
    "rtb_partition_shift": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _vp]),
    "rtb_partition_count_rows": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "rtb_partition_localize": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp]),
    "rtb_partition_max_ikey": (C.c_int, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, _vp]),
    "rtb_partition_compose_ikey": (C.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "rtb_valid_mask": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp]),
    "rtb_merge_row_counts": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "rtb_merge_scan_workspace_bytes": (_sz, [_i64]),
    "rtb_merge_scan_counts": (C.c_int, [_vp, _i64, _vp, _pi64, _vp, _sz, _vp]),
    "rtb_merge_expand": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rtb_reverse_shuffle": (C.c_int, [_vp, _i64, _vp, _vp]),
    "rtb_remap_ikey": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
}

MISSING = []
for _name, (_res, _args) in PROTOTYPES.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        MISSING.append(_name)
        continue
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.rtb_last_error().decode("utf-8", "replace")


def profile_get(pid: int):
    ms, launches, units = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
    lib.rtb_profile_get(pid, C.byref(ms), C.byref(launches), C.byref(units))
    return ms.value, launches.value, units.value


def launch_count() -> int:
    return int(lib.rtb_launch_count())


class RtbWorkspaceError(RuntimeError):
    def __init__(self, msg, needed):
        super().__init__(msg)
        self.needed = needed


def check(rc: int, needed: int = 0):
    """Map a C status to the Python exception riptide_cpp's callers expect."""
    if rc == RTB_OK:
        return
    msg = last_error()
    if rc == RTB_ERR_ARG:
        raise ValueError(msg)
    if rc == RTB_ERR_UNSUPPORTED:
        raise TypeError(msg)
    if rc == RTB_ERR_WORKSPACE:
        raise RtbWorkspaceError(msg, needed)
    raise RuntimeError(msg)
