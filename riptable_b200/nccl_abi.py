"""ctypes binding of libriptable_b200_nccl.so (include/riptable_b200_nccl.h): the multi-GPU C ABI, for callers that hold a raw
`ncclComm_t` (a C++ riptide_cpp binding; the tests).  `riptable_b200.dist` is the torch.distributed form of the same step.

Also binds the three libnccl calls needed to create a communicator without torch.distributed (ncclGetUniqueId,
ncclCommInitRank, ncclCommDestroy), so a test can hand a real `ncclComm_t` to the library.
"""
import ctypes as C
import os

from . import _lib  # loads libriptable_b200.so first (the NCCL library links against it)
from ._lib import rtb_column

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libriptable_b200_nccl.so")

RTB_SHARDED_LATE_KEYS = 1
NCCL_UNIQUE_ID_BYTES = 128


class rtb_sharded_result(C.Structure):
    _fields_ = [("unique_count", C.c_int64), ("late_keys", C.c_int64), ("seed_rows", C.c_int64)]


class ncclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * NCCL_UNIQUE_ID_BYTES)]


if not os.path.exists(LIB_PATH):
    raise ImportError(f"riptable_b200: {LIB_PATH} is missing. Build it with `python -m riptable_b200.build`.")

lib = C.CDLL(LIB_PATH)
PROTOTYPES = {
    "rtb_sharded_groupby_sum32_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int]),
    "rtb_sharded_groupby_sum32_nccl": (
        C.c_int,
        [C.c_void_p, C.c_int, C.c_int, C.POINTER(rtb_column), C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
         C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(rtb_sharded_result), C.c_void_p, C.c_size_t, C.c_void_p],
    ),
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

# libnccl is already in the process (DT_NEEDED of the library above; torch's own copy if torch was imported first)
nccl = C.CDLL("libnccl.so.2")
nccl.ncclGetUniqueId.restype = C.c_int
nccl.ncclGetUniqueId.argtypes = [C.POINTER(ncclUniqueId)]
nccl.ncclCommInitRank.restype = C.c_int
nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, ncclUniqueId, C.c_int]
nccl.ncclCommDestroy.restype = C.c_int
nccl.ncclCommDestroy.argtypes = [C.c_void_p]


def get_unique_id() -> bytes:
    uid = ncclUniqueId()
    rc = nccl.ncclGetUniqueId(C.byref(uid))
    if rc != 0:
        raise RuntimeError(f"ncclGetUniqueId failed: {rc}")
    return C.string_at(C.byref(uid), NCCL_UNIQUE_ID_BYTES)


def comm_init_rank(uid_bytes: bytes, nranks: int, rank: int) -> C.c_void_p:
    uid = ncclUniqueId()
    C.memmove(C.byref(uid), uid_bytes, NCCL_UNIQUE_ID_BYTES)
    comm = C.c_void_p()
    rc = nccl.ncclCommInitRank(C.byref(comm), nranks, uid, rank)
    if rc != 0:
        raise RuntimeError(f"ncclCommInitRank failed: {rc}")
    return comm


def comm_destroy(comm):
    nccl.ncclCommDestroy(comm)
