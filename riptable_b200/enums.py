"""Contract constants shared by both sides of the riptide_cpp boundary.

Restated (not copied) from the reference's enum module:
  * op numbers   -> riptable/rt_enum.py:486-532  (GB_FUNCTIONS)
  * sentinels    -> riptable/rt_enum.py:88-143   (INVALID_DICT, non-win32 table)
  * index dtype  -> riptable/rt_enum.py:666-679  (int_dtype_from_len)

The C side mirrors these in include/riptable_b200.h (rtb_dtype / rtb_gb_func).
"""
from enum import IntEnum

import numpy as np


class GB_FUNCTIONS(IntEnum):
    # unpacked per-bin reductions (rc.GroupByAll32)
    GB_SUM = 0
    GB_MEAN = 1
    GB_MIN = 2
    GB_MAX = 3
    GB_VAR = 4
    GB_STD = 5
    GB_NANSUM = 50
    GB_NANMEAN = 51
    GB_NANMIN = 52
    GB_NANMAX = 53
    GB_NANVAR = 54
    GB_NANSTD = 55
    # packed ops (rc.GroupByAllPack32)
    GB_FIRST = 100
    GB_NTH = 101
    GB_LAST = 102
    GB_MEDIAN = 103
    GB_MODE = 104
    GB_QUANTILE_MULT = 106
    # rolling ops on the packed layout (rc.GroupByAllPack32, row-shaped results)
    GB_ROLLING_SUM = 200
    GB_ROLLING_NANSUM = 201
    GB_ROLLING_DIFF = 202
    GB_ROLLING_SHIFT = 203
    GB_ROLLING_COUNT = 204
    GB_ROLLING_MEAN = 205
    GB_ROLLING_NANMEAN = 206
    GB_ROLLING_QUANTILE = 207
    # row-shaped accumulations (rc.EmaAll32)
    GB_CUMSUM = 300
    GB_EMADECAY = 301
    GB_EMANORMAL = 304
    GB_EMAWEIGHTED = 305
    GB_CUMPROD = 302
    GB_CUMNANMAX = 306
    GB_CUMNANMIN = 307
    GB_CUMMAX = 308
    GB_CUMMIN = 309


GB_FUNC_COUNT = -1

# Canonical element type codes used across the C ABI (include/riptable_b200.h: rtb_dtype).
RTB_BOOL, RTB_I8, RTB_U8, RTB_I16, RTB_U16, RTB_I32, RTB_U32, RTB_I64, RTB_U64, RTB_F32, RTB_F64, RTB_BYTES, RTB_UCS4 = range(13)

_KIND_SIZE_TO_CODE = {
    ("b", 1): RTB_BOOL,
    ("i", 1): RTB_I8,
    ("u", 1): RTB_U8,
    ("i", 2): RTB_I16,
    ("u", 2): RTB_U16,
    ("i", 4): RTB_I32,
    ("u", 4): RTB_U32,
    ("i", 8): RTB_I64,
    ("u", 8): RTB_U64,
    ("f", 4): RTB_F32,
    ("f", 8): RTB_F64,
}

_CODE_TO_NP = {
    RTB_BOOL: np.bool_,
    RTB_I8: np.int8,
    RTB_U8: np.uint8,
    RTB_I16: np.int16,
    RTB_U16: np.uint16,
    RTB_I32: np.int32,
    RTB_U32: np.uint32,
    RTB_I64: np.int64,
    RTB_U64: np.uint64,
    RTB_F32: np.float32,
    RTB_F64: np.float64,
}


def dtype_code(dt) -> int:
    """numpy dtype -> rtb_dtype code; fixed-width S/U/void items map to RTB_BYTES."""
    dt = np.dtype(dt)
    if dt.kind == "U":
        return RTB_UCS4
    if dt.kind in "SV":
        return RTB_BYTES
    if dt.kind in "mM":  # datetime64/timedelta64 are int64 payloads
        return RTB_I64
    try:
        return _KIND_SIZE_TO_CODE[(dt.kind, dt.itemsize)]
    except KeyError:
        raise TypeError(f"unsupported dtype {dt!r}") from None


def code_dtype(code: int):
    return np.dtype(_CODE_TO_NP[code])


def invalid_for(dt):
    """The dtype's invalid sentinel (rt_enum.py:116-143)."""
    dt = np.dtype(dt)
    if dt.kind == "f":
        return dt.type(np.nan)
    if dt.kind == "b":
        return np.bool_(False)
    if dt.kind == "i":
        return dt.type(np.iinfo(dt).min)
    if dt.kind == "u":
        return dt.type(np.iinfo(dt).max)
    if dt.kind == "S":
        return b""
    if dt.kind == "U":
        return ""
    raise TypeError(f"no invalid for {dt!r}")


def int_dtype_from_len(newlen: int) -> np.dtype:
    """Stored index dtype policy (rt_enum.py:666-679)."""
    if newlen < 100:
        return np.dtype(np.int8)
    if newlen < 30_000:
        return np.dtype(np.int16)
    if newlen < 2_000_000_000:
        return np.dtype(np.int32)
    return np.dtype(np.int64)


def sum_out_dtype(dt):
    """sum/nansum: ints and bool widen to int64, uint64 stays uint64, floats keep their dtype
    (rt_groupbynumba.py:636-648 states the intent; uint64 is unpinned)."""
    dt = np.dtype(dt)
    if dt.kind == "f":
        return dt
    if dt.kind == "u" and dt.itemsize == 8:
        return np.dtype(np.uint64)
    return np.dtype(np.int64)
