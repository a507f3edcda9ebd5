"""ctypes wrapper of oracle/liboracle.so (C restatement; test infrastructure / CPU baseline only)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_groupby_u64.restype = C.c_int64
        _lib.orc_groupby_u64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.orc_groupby_u64_parallel.restype = C.c_int64
        _lib.orc_groupby_u64_parallel.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_sum_f64.restype = None
        _lib.orc_sum_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int]
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def groupby_u64(keys, filter=None, threads=1):
    """threads = 1: one sequential pass (rc.MultiKeyGroupBy32 without cutoffs); threads > 1: partitions grouped in
    parallel and merged (riptable's cutoffs + hstack_groupings route).  Same result either way."""
    keys = np.ascontiguousarray(keys)
    assert keys.dtype.itemsize == 8
    n = len(keys)
    ikey = np.empty(n, np.int32)
    ifirst = np.empty(max(n, 1), np.int32)
    f = np.ascontiguousarray(filter, dtype=np.uint8) if filter is not None else None
    fp = f.ctypes.data if f is not None else None
    if threads and threads > 1:
        u = lib().orc_groupby_u64_parallel(keys.ctypes.data, fp, n, ikey.ctypes.data, ifirst.ctypes.data, int(threads))
    else:
        u = lib().orc_groupby_u64(keys.ctypes.data, fp, n, ikey.ctypes.data, ifirst.ctypes.data)
    if u < 0:
        raise MemoryError("orc_groupby_u64")
    return ikey, ifirst[:u].copy(), int(u)


def sum_f64(values, ikey, unique_rows, lo=1, hi=None, threads=None):
    values = np.ascontiguousarray(values, dtype=np.float64)
    ikey = np.ascontiguousarray(ikey, dtype=np.int32)
    nb = unique_rows + 1
    out = np.empty(nb, np.float64)
    lib().orc_sum_f64(values.ctypes.data, ikey.ctypes.data, len(ikey), nb, lo, nb if hi is None else hi, out.ctypes.data,
                      threads or max_threads())
    return out
