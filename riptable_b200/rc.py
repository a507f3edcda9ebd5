"""Drop-in mirror of the riptide_cpp (`rc`) entry points on riptable's groupby / reduce / lexsort path.

Same names, positional arguments, keyword names (`cutoffs=`, `pack=`, `index=`, `ascending=`,
`base_index=`) and error behaviour as the calls in riptable/rt_numpy.py:1463-2348, 2658-2673 and
riptable/rt_grouping.py:3427-3454, so `import riptable_b200.rc as rc` can stand in for
`import riptide_cpp as rc` on this path.

Every function is a thin shim: stage inputs on the GPU (torch is used only for device memory and the
current stream), call the C ABI of libriptable_b200.so, hand results back.

  host mode   : inputs are NumPy arrays  -> outputs are NumPy arrays (H2D / D2H inside the call)
  device mode : any input is a CUDA tensor / DevArray -> outputs stay on the device (torch tensors,
                DevArray for byte-string columns), so chained calls never touch PCIe.

There is no CPU fallback: without a CUDA device or the built library these functions raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np
import torch

from . import _lib
from ._lib import lib, rtb_column, check, RtbWorkspaceError
from .enums import (
    GB_FUNCTIONS,
    RTB_BYTES,
    RTB_I32,
    code_dtype,
    dtype_code,
)

_TORCH_OF = {
    np.dtype(np.bool_): torch.bool,
    np.dtype(np.int8): torch.int8,
    np.dtype(np.uint8): torch.uint8,
    np.dtype(np.int16): torch.int16,
    np.dtype(np.uint16): torch.uint16,
    np.dtype(np.int32): torch.int32,
    np.dtype(np.uint32): torch.uint32,
    np.dtype(np.int64): torch.int64,
    np.dtype(np.uint64): torch.uint64,
    np.dtype(np.float32): torch.float32,
    np.dtype(np.float64): torch.float64,
}
_NP_OF = {v: k for k, v in _TORCH_OF.items()}


_PINNED_D2H_MIN = 1 << 20


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("riptable_b200: no CUDA device; this path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DevArray:
    """A device-resident 1-D array of any NumPy dtype (incl. fixed-width S/U), backed by a uint8 CUDA tensor."""

    __slots__ = ("buf", "dtype", "n")

    def __init__(self, buf: torch.Tensor, dtype, n: int):
        self.buf, self.dtype, self.n = buf, np.dtype(dtype), int(n)

    def __len__(self):
        return self.n

    @property
    def ptr(self):
        return self.buf.data_ptr()

    def numpy(self) -> np.ndarray:
        nbytes = self.n * self.dtype.itemsize
        if nbytes == 0:
            return np.zeros(self.n, self.dtype)
        src = self.buf[:nbytes]
        if nbytes >= _PINNED_D2H_MIN:
            # large results land in page-locked memory (torch's caching host allocator), which the copy engine
            # fills at PCIe rate; the returned array owns the buffer
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            host.copy_(src, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            host = src.cpu()
        return host.numpy().view(self.dtype)

    def torch(self) -> torch.Tensor:
        t = self.buf[: self.n * self.dtype.itemsize]
        td = _TORCH_OF.get(self.dtype)
        if td is None:
            raise TypeError(f"{self.dtype} has no torch dtype; use .numpy() or .buf")
        return t.view(td)

    @staticmethod
    def empty(n, dtype, dev=None):
        dtype = np.dtype(dtype)
        nbytes = max(int(n) * dtype.itemsize, 16)
        return DevArray(torch.empty(nbytes, dtype=torch.uint8, device=dev or _device()), dtype, n)


def _is_device(x) -> bool:
    return isinstance(x, DevArray) or (isinstance(x, torch.Tensor) and x.is_cuda)


# ---- pageable host memory -> device -------------------------------------------------------------------------------
# A NumPy array riptable hands in lives in pageable memory; the driver's own staging of such a copy runs at about
# 10 GB/s.  Large pageable sources are therefore cut into chunks that worker threads copy into a small ring of
# page-locked buffers (NumPy releases the GIL for the memcpy) while the copy engine ships the previous chunks.
_STAGE_MIN = 8 << 20
_STAGE_CHUNK = 8 << 20
_STAGE_SLOTS = 6
_stage = None
_stage_lock = None


def _stage_ring():
    global _stage
    if _stage is None:
        from concurrent.futures import ThreadPoolExecutor

        bufs = [torch.empty(_STAGE_CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(_STAGE_SLOTS)]
        _stage = {"bufs": bufs, "np": [b.numpy() for b in bufs], "ev": [None] * _STAGE_SLOTS,
                  "pool": ThreadPoolExecutor(max_workers=4, thread_name_prefix="rtb-stage")}
    return _stage


def _h2d_staged(dst: torch.Tensor, src: np.ndarray):
    """dst (device uint8) <- src (pageable uint8 NumPy), chunked through the page-locked ring."""
    global _stage_lock
    if _stage_lock is None:
        import threading

        _stage_lock = threading.Lock()
    with _stage_lock:
        _h2d_staged_locked(dst, src)


def _h2d_staged_locked(dst, src):
    ring = _stage_ring()
    n = src.nbytes
    nchunks = (n + _STAGE_CHUNK - 1) // _STAGE_CHUNK
    stream = torch.cuda.current_stream()

    def fill(slot, lo, hi):
        ev = ring["ev"][slot]
        if ev is not None:
            ev.synchronize()  # the copy engine has drained this slot
        np.copyto(ring["np"][slot][: hi - lo], src[lo:hi])

    futs = {}
    for c in range(min(_STAGE_SLOTS, nchunks)):
        futs[c] = ring["pool"].submit(fill, c % _STAGE_SLOTS, c * _STAGE_CHUNK, min(n, (c + 1) * _STAGE_CHUNK))
    for c in range(nchunks):
        slot, lo, hi = c % _STAGE_SLOTS, c * _STAGE_CHUNK, min(n, (c + 1) * _STAGE_CHUNK)
        futs.pop(c).result()
        dst[lo:hi].copy_(ring["bufs"][slot][: hi - lo], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        ring["ev"][slot] = ev
        nxt = c + _STAGE_SLOTS
        if nxt < nchunks:
            futs[nxt] = ring["pool"].submit(fill, slot, nxt * _STAGE_CHUNK, min(n, (nxt + 1) * _STAGE_CHUNK))


def _to_dev(x) -> DevArray:
    """Stage one 1-D array on the GPU as raw bytes (zero-copy for CUDA tensors)."""
    if isinstance(x, DevArray):
        return x
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            if x.dim() != 1:
                raise ValueError("expected a 1-D tensor")
            x = x.contiguous()
            if x.data_ptr() % 16:
                x = x.clone()
            npdt = _NP_OF[x.dtype]
            return DevArray(x.view(torch.uint8) if x.numel() else torch.empty(16, dtype=torch.uint8, device=x.device), npdt, x.numel())
        x = x.numpy()
    a = np.asarray(x)
    if a.ndim != 1:
        raise ValueError("expected a 1-D array")
    if a.dtype.kind == "O":
        raise TypeError("object arrays are not supported on this path")
    if not a.dtype.isnative:
        a = a.astype(a.dtype.newbyteorder("="))
    a = np.ascontiguousarray(a)
    out = DevArray.empty(len(a), a.dtype)
    nbytes = a.nbytes
    if nbytes:
        raw = a.view(np.uint8).reshape(-1)
        src = torch.from_numpy(raw)
        if src.is_pinned():
            out.buf[:nbytes].copy_(src, non_blocking=True)
        elif nbytes >= _STAGE_MIN and os.environ.get("RTB_NO_STAGING") != "1":
            _h2d_staged(out.buf, raw)
        else:
            out.buf[:nbytes].copy_(src)
    return out


def _from_dev(d: DevArray, device_mode: bool):
    if not device_mode:
        return d.numpy()
    if d.dtype in _TORCH_OF:
        return d.torch()
    return d


def _col(d: DevArray) -> rtb_column:
    code = dtype_code(d.dtype)
    return rtb_column(C.c_void_p(d.ptr), code, d.dtype.itemsize)


def _cols(devs: Sequence[DevArray]):
    arr = (rtb_column * len(devs))()
    for i, d in enumerate(devs):
        arr[i] = _col(d)
    return arr


def _ikey_dev(ikey) -> DevArray:
    d = _to_dev(ikey)
    if d.dtype.kind != "i":
        raise ValueError("ikey must be int8, int16, int32, or int64")
    return d


def _bool_dev(f, n, what="filter") -> DevArray:
    d = _to_dev(f)
    if d.dtype.kind != "b" and d.dtype != np.dtype(np.uint8):
        raise TypeError(f"{what} must be a boolean array")
    if d.n != n:
        raise ValueError(f"{what} length {d.n} does not match {n}")
    return d


def _ws(nbytes: int) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=_device())


def _cutoffs_dev(cutoffs, n):
    """cutoffs (int64 cumulative row ends) validated on the host and staged on the device, with the partition id of
    every row.  -> (host int64 array, device int64 tensor, pid int32 tensor[n])"""
    cut = np.ascontiguousarray(np.asarray(cutoffs.cpu() if isinstance(cutoffs, torch.Tensor) else cutoffs), dtype=np.int64)
    if cut.ndim != 1 or len(cut) == 0 or cut[-1] != n or cut[0] < 0 or (np.diff(cut) < 0).any():
        raise ValueError("cutoffs must be a non-decreasing 1-D int64 array whose last entry is the row count")
    dev = _device()
    cdev = torch.from_numpy(cut).to(dev)
    pid = torch.empty(max(n, 4), dtype=torch.int32, device=dev)
    check(lib.rtb_partition_ids(n, C.c_void_p(cdev.data_ptr()), len(cut), C.c_void_p(pid.data_ptr()), _stream()))
    return cut, cdev, pid[:n]


def _as_list(list_arrays):
    if isinstance(list_arrays, (np.ndarray, torch.Tensor, DevArray)):
        return [list_arrays]
    if isinstance(list_arrays, tuple):
        return list(list_arrays)
    return list_arrays


def _initial_max_unique(n: int, hint_size) -> int:
    """Bins the first attempt provisions for.  `hint_size` (riptable's own estimate of the unique count,
    riptable/rt_numpy.py:2066-2070) is taken as given; without one, room for n/8 uniques (at least 1 Mi).  The C side
    reports RTB_ERR_WORKSPACE with an extrapolated requirement when a call needs more, and the shim retries with it, so
    a 1 B-row call with 1 M keys no longer reserves a table region sized for 500 M."""
    if hint_size and hint_size > 0:
        return min(n, max(int(hint_size), 1024))
    return min(n, max(1 << 20, n // 8))


# ---- host mode, large inputs: PCIe in both directions at once ---------------------------------------------------------
# A host-mode call moves its inputs up and its iKey down the same link.  The link is full duplex (measured on this pool:
# 55 GB/s up alone, 57 down alone, 99 GB/s with both directions busy, profiles/r02/pcie_duplex.py), so the rows are cut into
# chunks that go through the resumable grouping (GroupByStream): while chunk i is grouped, chunk i + 1 .. i + 2 are on
# their way up and the iKey of chunk i - 1 is on its way down.  Results are those of the one-shot call (the resumable
# grouping numbers bins as if the chunks were one array).
_PIPE_MIN_ROWS = 48 << 20
_PIPE_CHUNK_ROWS = int(os.environ.get("RTB_PIPE_CHUNK_ROWS", 32 << 20)) & ~1023  # measurement override
_PIPE_DEPTH = 3


def _pipeline_ok(list_arrays, values, filter):
    if os.environ.get("RTB_NO_PIPELINE") == "1" or len(list_arrays) != 1:
        return False
    k = list_arrays[0]
    if not isinstance(k, np.ndarray) or k.ndim != 1 or len(k) < _PIPE_MIN_ROWS or len(k) > 2_147_483_646:
        return False
    if k.dtype.kind not in "biuf" or k.dtype.itemsize not in (1, 2, 4, 8) or k.dtype == np.float16 or not k.dtype.isnative:
        return False
    if not k.flags.c_contiguous:
        return False
    if values is not None:
        if not isinstance(values, np.ndarray) or values.ndim != 1 or len(values) != len(k) or not values.flags.c_contiguous:
            return False
        if values.dtype.kind not in "biuf" or values.dtype == np.float16 or not values.dtype.isnative:
            return False
    if filter is not None:
        if not isinstance(filter, np.ndarray) or filter.dtype != np.bool_ or len(filter) != len(k) or not filter.flags.c_contiguous:
            return False
    return True


def _h2d_chunk(dst: torch.Tensor, src: np.ndarray, stream):
    """dst (device uint8) <- the bytes of src, enqueued on `stream`."""
    raw = src.view(np.uint8).reshape(-1)
    t = torch.from_numpy(raw)
    with torch.cuda.stream(stream):
        if t.is_pinned():
            dst[: raw.nbytes].copy_(t, non_blocking=True)
        elif raw.nbytes >= _STAGE_MIN and os.environ.get("RTB_NO_STAGING") != "1":
            _h2d_staged(dst[: raw.nbytes], raw)
        else:
            dst[: raw.nbytes].copy_(t)


def _groupby_pipelined(key: np.ndarray, values, filter, hint_size, func, bin_low):
    """Chunk-pipelined host-mode MultiKeyGroupBy32 / MultiKeyGroupBySum32 for one numeric key column.
    -> (iKey ndarray, iFirstKey ndarray, unique_count, sums ndarray | None), or None if the provisioned table overflowed."""
    dev = _device()
    n = len(key)
    C_ = _PIPE_CHUNK_ROWS
    nchunks = (n + C_ - 1) // C_
    max_unique = _initial_max_unique(n, hint_size)
    with_sums = values is not None
    st = GroupByStream(max_unique, C_, with_sums=with_sums, func=func, bin_low=bin_low)
    ksz = key.dtype.itemsize
    vsz = values.dtype.itemsize if with_sums else 0
    cur = torch.cuda.current_stream()
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    kbuf = [torch.empty(C_ * ksz, dtype=torch.uint8, device=dev) for _ in range(_PIPE_DEPTH)]
    vbuf = [torch.empty(C_ * vsz, dtype=torch.uint8, device=dev) for _ in range(_PIPE_DEPTH)] if with_sums else None
    fbuf = [torch.empty(C_, dtype=torch.uint8, device=dev) for _ in range(_PIPE_DEPTH)] if filter is not None else None
    ibuf = [torch.empty(C_, dtype=torch.int32, device=dev) for _ in range(_PIPE_DEPTH)]
    up_ev = [None] * _PIPE_DEPTH
    down_ev = [None] * _PIPE_DEPTH
    host_ikey = torch.empty(n, dtype=torch.int32, pin_memory=True)
    up.wait_stream(cur)

    def send(c):
        s_ = c % _PIPE_DEPTH
        lo, hi = c * C_, min(n, (c + 1) * C_)
        _h2d_chunk(kbuf[s_], key[lo:hi], up)
        if with_sums:
            _h2d_chunk(vbuf[s_], values[lo:hi], up)
        if filter is not None:
            _h2d_chunk(fbuf[s_], filter[lo:hi], up)
        ev = torch.cuda.Event()
        ev.record(up)
        up_ev[s_] = ev

    for c in range(min(_PIPE_DEPTH - 1, nchunks)):
        send(c)
    try:
        for c in range(nchunks):
            s_ = c % _PIPE_DEPTH
            lo, hi = c * C_, min(n, (c + 1) * C_)
            m = hi - lo
            if c + _PIPE_DEPTH - 1 < nchunks:
                send(c + _PIPE_DEPTH - 1)  # its slot was consumed by the (synchronous) push of chunk c - 1
            cur.wait_event(up_ev[s_])
            if down_ev[s_] is not None:
                down_ev[s_].synchronize()  # the iKey buffer of this slot has been read back
            kd = DevArray(kbuf[s_], key.dtype, m)
            vd = DevArray(vbuf[s_], values.dtype, m) if with_sums else None
            fd = fbuf[s_][:m].view(torch.bool) if filter is not None else None
            st.push(kd, fd, lo, ibuf[s_], values=vd) if with_sums else st.push(kd, fd, lo, ibuf[s_])
            down.wait_stream(cur)
            with torch.cuda.stream(down):
                host_ikey[lo:hi].copy_(ibuf[s_][:m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(down)
            down_ev[s_] = ev
    except RtbWorkspaceError:
        up.synchronize()
        down.synchronize()
        return None
    u = st.unique_count
    ifk = st.ifirstkey.cpu().numpy().copy()
    sums = None
    if with_sums:
        raw = st.sums()
        odt = code_dtype(lib.rtb_reduce_out_dtype(int(func), dtype_code(values.dtype)))
        if odt == np.dtype(np.float32):
            sums = raw.to(torch.float32).cpu().numpy()
        elif odt == np.dtype(np.uint64):
            sums = raw.cpu().numpy().view(np.uint64)
        else:
            sums = raw.cpu().numpy()
    down.synchronize()
    cur.wait_stream(up)
    return host_ikey.numpy(), ifk, u, sums


# ==================================================================================================
# a1  rc.MultiKeyGroupBy32 (riptable/rt_numpy.py:2073-2075)
def MultiKeyGroupBy32(list_arrays, hint_size=0, filter=None, hash_mode=2, cutoffs=None):
    """-> (iKey int32[N] base-1 / 0 = filtered, iFirstKey int32[U], unique_count).
    With `cutoffs` numbering restarts per partition and iFirstKey is (rows, cumulative unique cutoffs)
    (riptable/tests/test_pgroupby.py:7-86)."""
    list_arrays = _as_list(list_arrays)
    if not isinstance(list_arrays, list) or len(list_arrays) == 0:
        raise ValueError("groupbyhash first argument is not a list of numpy arrays")
    if cutoffs is None and _pipeline_ok(list_arrays, None, filter):
        res = _groupby_pipelined(list_arrays[0], None, filter, hint_size, GB_FUNCTIONS.GB_SUM, 1)
        if res is not None:
            return res[0], res[1], res[2]
    device_mode = any(_is_device(a) for a in list_arrays)
    cols = [_to_dev(a) for a in list_arrays]
    n = cols[0].n
    if any(c.n != n for c in cols):
        raise ValueError(f"groupbyhash all arrays must have same length not {set(c.n for c in cols)}")
    fdev = _bool_dev(filter, n) if filter is not None else None
    cut = None
    if cutoffs is not None:
        cut = np.ascontiguousarray(np.asarray(cutoffs.cpu() if isinstance(cutoffs, torch.Tensor) else cutoffs), dtype=np.int64)
        if cut.ndim != 1 or len(cut) == 0 or cut[-1] != n:
            raise ValueError("cutoffs must be a 1-D int64 array whose last entry is the row count")
    dev = _device()
    if n == 0:
        e = DevArray.empty(0, np.int32)
        return _from_dev(e, device_mode), _from_dev(e, device_mode), 0

    max_unique = _initial_max_unique(n, hint_size)
    carr = _cols(cols)
    ikey = DevArray.empty(n, np.int32, dev)
    ucount, needed = C.c_int64(0), C.c_int64(0)
    ucut = np.zeros(len(cut), dtype=np.int64) if cut is not None else None
    while True:
        ifirst = DevArray.empty(max_unique, np.int32, dev)
        direct = len(cols) == 1 and cols[0].dtype.kind in "biuf" and cut is None
        row_bytes = 0 if direct else sum(c.dtype.itemsize for c in cols)
        ws = _ws(lib.rtb_groupby_workspace_bytes_rows(n, max_unique, row_bytes, int(cut is not None)))
        rc_ = lib.rtb_multikey_groupby32(
            carr, len(cols), n, int(hint_size or 0),
            C.c_void_p(fdev.ptr) if fdev is not None else None,
            cut.ctypes.data_as(C.POINTER(C.c_int64)) if cut is not None else None, len(cut) if cut is not None else 0,
            C.c_void_p(ikey.ptr), C.c_void_p(ifirst.ptr), max_unique,
            C.byref(ucount), ucut.ctypes.data_as(C.POINTER(C.c_int64)) if ucut is not None else None, C.byref(needed),
            C.c_void_p(ws.data_ptr()), ws.numel(), _stream(),
        )
        if rc_ == _lib.RTB_ERR_WORKSPACE and max_unique < n:
            max_unique = min(n, max(int(needed.value), 4 * max_unique))
            continue
        check(rc_, needed.value)
        break
    u = int(ucount.value)
    ifirst = DevArray(ifirst.buf, np.int32, u)
    ik, ifk = _from_dev(ikey, device_mode), _from_dev(ifirst, device_mode)
    if device_mode and u < max_unique // 2 and max_unique > (4 << 20):
        ifk = ifk.clone()  # do not pin a large max_unique-sized buffer behind a small result
    if cut is not None:
        return ik, (ifk, ucut), u
    return ik, ifk, u


# ==================================================================================================
# a1 + a5 fused: what `ds.gb(key).sum()` / `Categorical(key).sum(v)` run back to back (riptable/rt_groupbyops.py:1161-1225)
def MultiKeyGroupBySum32(list_arrays, values, hint_size=0, filter=None, hash_mode=2, func=GB_FUNCTIONS.GB_SUM, bin_low=1):
    """MultiKeyGroupBy32(list_arrays, hint_size, filter) and GroupByAll32([values], iKey, U, [func], [bin_low], [U + 1])
    in one pass over the rows.  func: GB_SUM or GB_NANSUM.
    -> (iKey int32[N], iFirstKey int32[U], unique_count, sums[U + 1]) with exactly the values of the two calls."""
    list_arrays = _as_list(list_arrays)
    if not isinstance(list_arrays, list) or len(list_arrays) == 0:
        raise ValueError("groupbyhash first argument is not a list of numpy arrays")
    if int(func) in (GB_FUNCTIONS.GB_SUM, GB_FUNCTIONS.GB_NANSUM) and isinstance(values, np.ndarray) and _pipeline_ok(list_arrays, values, filter):
        res = _groupby_pipelined(list_arrays[0], values, filter, hint_size, int(func), int(bin_low))
        if res is not None:
            return res
    device_mode = any(_is_device(a) for a in list_arrays) or _is_device(values)
    cols = [_to_dev(a) for a in list_arrays]
    n = cols[0].n
    if any(c.n != n for c in cols):
        raise ValueError(f"groupbyhash all arrays must have same length not {set(c.n for c in cols)}")
    v = _to_dev(values)
    if v.n != n:
        raise ValueError(f"GroupByAll32: value length {v.n} does not match ikey length {n}")
    if v.dtype.kind not in "biuf" or v.dtype == np.float16:
        raise TypeError(f"GroupByAll32: {v.dtype} cannot be summed")
    vcode = dtype_code(v.dtype)
    ocode = lib.rtb_reduce_out_dtype(int(func), vcode)
    if ocode < 0 or int(func) not in (GB_FUNCTIONS.GB_SUM, GB_FUNCTIONS.GB_NANSUM):
        raise TypeError("MultiKeyGroupBySum32: GB_SUM / GB_NANSUM of a numeric column")
    odt = code_dtype(ocode)
    fdev = _bool_dev(filter, n) if filter is not None else None
    dev = _device()
    if n == 0:
        e = DevArray.empty(0, np.int32)
        z = DevArray(torch.zeros(16, dtype=torch.uint8, device=dev), odt, 1)
        return _from_dev(e, device_mode), _from_dev(e, device_mode), 0, _from_dev(z, device_mode)
    max_unique = _initial_max_unique(n, hint_size)
    carr = _cols(cols)
    ikey = DevArray.empty(n, np.int32, dev)
    single = len(cols) == 1 and cols[0].dtype.kind in "biuf"
    row_bytes = 0 if single else sum(c.dtype.itemsize for c in cols)
    ucount, needed = C.c_int64(0), C.c_int64(0)
    while True:
        ifirst = DevArray.empty(max_unique, np.int32, dev)
        sums = DevArray.empty(max_unique + 1, odt, dev)
        ws = _ws(lib.rtb_groupby_sum_workspace_bytes(n, max_unique, row_bytes))
        rc_ = lib.rtb_multikey_groupby_sum32(
            carr, len(cols), n, C.c_void_p(fdev.ptr) if fdev is not None else None, C.c_void_p(v.ptr), vcode, int(func),
            int(bin_low), C.c_void_p(ikey.ptr), C.c_void_p(ifirst.ptr), max_unique, C.c_void_p(sums.ptr), C.byref(ucount),
            C.byref(needed), C.c_void_p(ws.data_ptr()), ws.numel(), _stream())
        if rc_ == _lib.RTB_ERR_WORKSPACE and max_unique < n:
            max_unique = min(n, max(int(needed.value), 4 * max_unique))
            continue
        check(rc_, needed.value)
        break
    u = int(ucount.value)
    ifirst = DevArray(ifirst.buf, np.int32, u)
    sums = DevArray(sums.buf, odt, u + 1)
    ik, ifk, sm = _from_dev(ikey, device_mode), _from_dev(ifirst, device_mode), _from_dev(sums, device_mode)
    if device_mode and u < max_unique // 2 and max_unique > (4 << 20):
        ifk, sm = ifk.clone(), sm.clone()  # do not pin large max_unique-sized buffers behind small results
    return ik, ifk, u, sm


# ==================================================================================================
# a4  rc.CombineFilter (riptable/rt_numpy.py:1508)
def CombineFilter(key, filter):
    device_mode = _is_device(key) or _is_device(filter)
    k = _ikey_dev(key)
    f = _bool_dev(filter, k.n)
    out = DevArray.empty(k.n, k.dtype)
    check(lib.rtb_combine_filter(C.c_void_p(k.ptr), dtype_code(k.dtype), C.c_void_p(f.ptr), k.n, C.c_void_p(out.ptr), _stream()))
    return _from_dev(out, device_mode)


def CombineAccum1Filter(key1, unique_count1, filter=None):
    """rc.CombineAccum1Filter (riptable/rt_numpy.py:1512-1545): drop filtered rows and bin 0, renumber the surviving
    bins by first appearance.  -> (iKey in key1's dtype, iFirstKey int32, unique_count).  The renumbering is the
    hash grouping itself, applied to the bin ids."""
    device_mode = _is_device(key1) or _is_device(filter)
    k = _ikey_dev(key1)
    if filter is not None:
        k = _combine_filter_dev(k, _bool_dev(filter, k.n))
    keep = torch.empty(max(k.n, 4), dtype=torch.uint8, device=_device())[: k.n]
    check(lib.rtb_positive_mask(C.c_void_p(k.ptr), dtype_code(k.dtype), k.n, C.c_void_p(keep.data_ptr()), _stream()))
    ik, ifk, u = MultiKeyGroupBy32([k], 0, keep.view(torch.bool), 2)
    ik = ik.to(_TORCH_OF[k.dtype]) if isinstance(ik, torch.Tensor) else ik
    if device_mode:
        return ik, ifk, u
    return ik.cpu().numpy(), ifk.cpu().numpy(), u


class GroupByStream:
    """Resumable MultiKeyGroupBy32 over consecutive row chunks of ONE 1/2/4/8-byte key column (device mode).

    `push(chunk)` returns the chunk's iKey; bins keep counting in first-appearance order across chunks, so the
    concatenation of the returned iKeys equals one MultiKeyGroupBy32 call on the concatenated keys.  Used by
    riptable_b200.dist (seeded shards) and by the host-mode pipeline that overlaps PCIe copies with hashing.

    With `with_sums=True` the stream is the resumable form of MultiKeyGroupBySum32: `push(chunk, values=v)` also adds
    every row's value to its bin (`sums()`), in the same pass; a push without values only creates bins."""

    def __init__(self, max_unique: int, max_chunk_rows: int, with_sums: bool = False, func=GB_FUNCTIONS.GB_SUM, bin_low: int = 1):
        self.max_unique = int(max_unique)
        self.max_chunk_rows = int(max_chunk_rows)
        self.state = _lib.rtb_gb_state()
        self._started = False
        self.ifirst = DevArray.empty(self.max_unique, np.int32)
        self.with_sums, self.func, self.bin_low = bool(with_sums), int(func), int(bin_low)
        self._vdtype = None
        if self.with_sums:
            self._acc = torch.empty(self.max_unique + 1, dtype=torch.int64, device=_device())
            self._ws = _ws(lib.rtb_groupby_sum_chunk_workspace_bytes(self.max_chunk_rows, self.max_unique))
        else:
            self._ws = _ws(lib.rtb_groupby_workspace_bytes(self.max_chunk_rows, self.max_unique, 1, 0))

    @property
    def unique_count(self) -> int:
        return int(self.state.unique)

    @property
    def ifirstkey(self) -> torch.Tensor:
        return self.ifirst.torch()[: self.unique_count]

    def sums(self) -> torch.Tensor:
        """Raw accumulators [unique_count + 1]: float64 for float value columns, int64 for integer ones."""
        a = self._acc[: self.unique_count + 1]
        return a.view(torch.float64) if self._vdtype is not None and self._vdtype.kind == "f" else a

    def adopt_seed(self, seed_keys):
        """Re-label this grouping with a seed, in place (rtb_multikey_groupby32_adopt_seed): `seed_keys` are distinct keys
        whose bins become 1..len(seed_keys); local keys the seed lacks follow in their first-appearance order.
        -> int32 CUDA tensor map[old bin] = new bin (re-index already returned iKey with it), or None when the table has
        no room for the seed (build a fresh stream from the seed instead)."""
        if not self._started:
            return None
        sk = _to_dev(seed_keys)
        u_loc = self.unique_count
        mp = torch.empty(u_loc + 4, dtype=torch.int32, device=_device())
        nlate = C.c_int64(0)
        col = _col(sk)
        rc_ = lib.rtb_multikey_groupby32_adopt_seed(
            C.byref(col), sk.n, C.c_void_p(self.ifirst.ptr), self.max_unique,
            C.c_void_p(self._acc.data_ptr()) if self.with_sums else None, C.byref(self.state), C.c_void_p(mp.data_ptr()),
            C.byref(nlate), C.c_void_p(self._ws.data_ptr()), self._ws.numel(), _stream())
        if rc_ == _lib.RTB_ERR_UNSUPPORTED:
            return None
        check(rc_)
        return mp[: u_loc + 1]

    def push(self, keys, filter=None, row_offset: int = 0, out=None, values=None):
        k = _to_dev(keys)
        if k.n > self.max_chunk_rows:
            raise ValueError(f"chunk of {k.n} rows exceeds max_chunk_rows={self.max_chunk_rows}")
        f = _bool_dev(filter, k.n) if filter is not None else None
        ik = out if out is not None else torch.empty(max(k.n, 4), dtype=torch.int32, device=_device())
        needed = C.c_int64(0)
        col = _col(k)
        if self.with_sums:
            v, vcode = None, dtype_code(self._vdtype) if self._vdtype is not None else RTB_I32
            if values is not None:
                v = _to_dev(values)
                if v.n != k.n:
                    raise ValueError("GroupByStream.push: values and keys differ in length")
                if v.dtype.kind not in "biuf" or v.dtype == np.float16:
                    raise TypeError(f"GroupByStream.push: {v.dtype} cannot be summed")
                if self._vdtype is not None and (self._vdtype.kind == "f") != (v.dtype.kind == "f"):
                    raise TypeError("GroupByStream.push: float and integer value chunks cannot share accumulators")
                self._vdtype, vcode = v.dtype, dtype_code(v.dtype)
            rc_ = lib.rtb_multikey_groupby_sum32_chunk(
                C.byref(col), k.n, C.c_void_p(f.ptr) if f is not None else None, C.c_void_p(v.ptr) if v is not None else None,
                vcode, self.func, self.bin_low, C.c_void_p(ik.data_ptr()), C.c_void_p(self.ifirst.ptr), self.max_unique,
                C.c_void_p(self._acc.data_ptr()), int(row_offset), C.byref(self.state), int(self._started), C.byref(needed),
                C.c_void_p(self._ws.data_ptr()), self._ws.numel(), _stream())
        else:
            if values is not None:
                raise ValueError("GroupByStream.push: create the stream with with_sums=True to sum values")
            rc_ = lib.rtb_multikey_groupby32_chunk(
                C.byref(col), k.n, C.c_void_p(f.ptr) if f is not None else None, C.c_void_p(ik.data_ptr()),
                C.c_void_p(self.ifirst.ptr), self.max_unique, int(row_offset), C.byref(self.state), int(self._started),
                C.byref(needed), C.c_void_p(self._ws.data_ptr()), self._ws.numel(), _stream())
        check(rc_, needed.value)
        self._started = True
        return ik[: k.n]


def _combine_filter_dev(k: DevArray, f: DevArray) -> DevArray:
    out = DevArray.empty(k.n, k.dtype)
    check(lib.rtb_combine_filter(C.c_void_p(k.ptr), dtype_code(k.dtype), C.c_void_p(f.ptr), k.n, C.c_void_p(out.ptr), _stream()))
    return out


def CombineAccum2Filter(key1, key2, unique_count1, unique_count2, filter=None):
    """rc.CombineAccum2Filter (riptable/rt_numpy.py:1549-1597) -> (iKey over the (unique_count1+1) x (unique_count2+1)
    cell grid, nCountGroup).  iKey = key2 * (unique_count1 + 1) + key1, 0 where the filter is False; its dtype follows
    riptable/rt_enum.py:666-679."""
    from .enums import int_dtype_from_len

    device_mode = any(_is_device(x) for x in (key1, key2, filter))
    k1, k2 = _ikey_dev(key1), _ikey_dev(key2)
    if k1.n != k2.n:
        raise ValueError("CombineAccum2Filter: keys must have the same length")
    f = _bool_dev(filter, k1.n) if filter is not None else None
    ncell = (int(unique_count1) + 1) * (int(unique_count2) + 1)
    odt = int_dtype_from_len(ncell)
    out = DevArray.empty(k1.n, odt)
    check(lib.rtb_combine_accum2(C.c_void_p(k1.ptr), dtype_code(k1.dtype), C.c_void_p(k2.ptr), dtype_code(k2.dtype),
                                 C.c_void_p(f.ptr) if f is not None else None, k1.n, int(unique_count1), C.c_void_p(out.ptr),
                                 dtype_code(odt), _stream()))
    cnt = BinCount(out, ncell) if device_mode else None
    if device_mode:
        return _from_dev(out, True), cnt
    cnt_dev = DevArray.empty(ncell, np.int32)
    check(lib.rtb_bincount(C.c_void_p(out.ptr), dtype_code(odt), out.n, ncell, C.c_void_p(cnt_dev.ptr), _stream()))
    return out.numpy(), cnt_dev.numpy()


# ==================================================================================================
# a5  rc.GroupByAll32 (riptable/rt_numpy.py:1871; riptable/rt_grouping.py:3434-3436)
def GroupByAll32(values, ikey, unique_rows, funcList, binLowList, binHighList, func_param=0):
    """-> tuple of arrays[unique_rows+1] (slot 0 = filtered bin); None for columns the op cannot take
    (riptable/rt_grouping.py:2341-2345)."""
    values = _as_list(values)
    device_mode = _is_device(ikey) or any(_is_device(v) for v in values)
    k = _ikey_dev(ikey)
    kcode = dtype_code(k.dtype)
    nb = int(unique_rows) + 1
    out = []
    for v, f, lo, hi in zip(values, funcList, binLowList, binHighList):
        d = _to_dev(v)
        if d.n != k.n:
            raise ValueError(f"GroupByAll32: value length {d.n} does not match ikey length {k.n}")
        vcode = dtype_code(d.dtype) if d.dtype.kind in "biuf" and d.dtype.itemsize <= 8 and d.dtype != np.float16 else RTB_BYTES
        ocode = lib.rtb_reduce_out_dtype(int(f), vcode) if vcode < RTB_BYTES else -1
        if ocode < 0:
            out.append(None)
            continue
        res = DevArray.empty(nb, code_dtype(ocode))
        ws = _ws(lib.rtb_groupby_reduce_workspace_bytes(int(unique_rows), int(f), vcode))
        check(lib.rtb_groupby_reduce(
            C.c_void_p(d.ptr), vcode, C.c_void_p(k.ptr), kcode, k.n, int(unique_rows), int(f), int(lo), int(hi),
            C.c_void_p(res.ptr), C.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
        out.append(_from_dev(res, device_mode))
    return tuple(out)


# ==================================================================================================
# a6 / a7  rc.BinCount / rc.GroupFromBinCount / rc.GroupByPack32 (riptable/rt_numpy.py:1945-1972)
def BinCount(ikey, n, pack=False):
    """histogram int32[n]; pack=True -> (nCountGroup, iGroup, iFirstGroup) (riptable/rt_numpy.py:1960)."""
    device_mode = _is_device(ikey)
    k = _ikey_dev(ikey)
    n = int(n)
    if pack:
        ig, ifg, cnt = _pack(k, n)
        return _from_dev(cnt, device_mode), _from_dev(ig, device_mode), _from_dev(ifg, device_mode)
    out = DevArray.empty(n, np.int32)
    check(lib.rtb_bincount(C.c_void_p(k.ptr), dtype_code(k.dtype), k.n, n, C.c_void_p(out.ptr), _stream()))
    return _from_dev(out, device_mode)


def _pack(k: DevArray, nbins: int):
    if k.n > 2_000_000_000:
        raise ValueError("pack: more than 2e9 rows needs int64 iGroup, which this build does not provide (iKey itself is int32-indexed)")
    ig = DevArray.empty(k.n, np.int32)
    ifg = DevArray.empty(nbins, np.int32)
    cnt = DevArray.empty(nbins, np.int32)
    ws = _ws(lib.rtb_pack_workspace_bytes(k.n, nbins))
    check(lib.rtb_groupby_pack(C.c_void_p(k.ptr), dtype_code(k.dtype), k.n, nbins, C.c_void_p(ig.ptr), C.c_void_p(ifg.ptr),
                               C.c_void_p(cnt.ptr), C.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
    return ig, ifg, cnt


def GroupFromBinCount(ikey, nCountGroup):
    """-> (iGroup, iFirstGroup) (riptable/rt_numpy.py:1957,1965)."""
    device_mode = _is_device(ikey) or _is_device(nCountGroup)
    k = _ikey_dev(ikey)
    ig, ifg, _ = _pack(k, len(nCountGroup))
    return _from_dev(ig, device_mode), _from_dev(ifg, device_mode)


def GroupByPack32(ikey, _unused, n, cutoffs=None):
    """-> (iGroup, iFirstGroup, nCountGroup) (riptable/rt_numpy.py:1972)."""
    device_mode = _is_device(ikey)
    if cutoffs is None:
        ig, ifg, cnt = _pack(_ikey_dev(ikey), int(n))
        return _from_dev(ig, device_mode), _from_dev(ifg, device_mode), _from_dev(cnt, device_mode)
    # cutoffs: every partition is packed as an array of its own (iKey numbered per partition, as groupbyhash(cutoffs=)
    # numbers it): U_p + 1 bins, U_p = its largest iKey.  iGroup holds partition-relative rows, partition after
    # partition; iFirstGroup / nCountGroup are the per-partition arrays laid end to end.  Unpinned by the reference's
    # tests, derived from the pinned hash case (oracle.GroupByPack32).  One pack over composite bins does it all.
    k = _ikey_dev(ikey)
    cut, cdev, pid = _cutoffs_dev(cutoffs, k.n)
    P = len(cut)
    dev = _device()
    umax = torch.empty(P, dtype=torch.int32, device=dev)
    bad = torch.empty(1, dtype=torch.int32, device=dev)
    check(lib.rtb_partition_max_ikey(C.c_void_p(k.ptr), dtype_code(k.dtype), k.n, C.c_void_p(pid.data_ptr()), P,
                                     C.c_void_p(umax.data_ptr()), C.c_void_p(bad.data_ptr()), _stream()))
    if int(bad.item()):
        raise ValueError("GroupByPack32: negative ikey values")
    nb_p = umax.to(torch.int64) + 1                      # bins of each partition, its slot 0 included
    off_incl = torch.cumsum(nb_p, 0)
    bin_off = (off_incl - nb_p).contiguous()
    nbins = int(off_incl[-1].item())
    if nbins >= 2**31:
        raise ValueError("GroupByPack32: too many bins")
    comp = DevArray.empty(k.n, np.int32)
    check(lib.rtb_partition_compose_ikey(C.c_void_p(k.ptr), dtype_code(k.dtype), k.n, C.c_void_p(pid.data_ptr()),
                                         C.c_void_p(bin_off.data_ptr()), C.c_void_p(comp.ptr), _stream()))
    ig, ifg, cnt = _pack(comp, nbins)
    if k.n:
        check(lib.rtb_partition_shift(C.c_void_p(ig.ptr), k.n, C.c_void_p(pid.data_ptr()), C.c_void_p(cdev.data_ptr()), P, 1, _stream()))
    # positions -> partition-relative positions (a table of `nbins` entries)
    starts = torch.cat([cdev.new_zeros(1), cdev[:-1]])
    part_of_bin = torch.searchsorted(off_incl, torch.arange(nbins, device=dev, dtype=torch.int64), right=True)
    ifg_t = ifg.torch()
    ifg_t -= starts[part_of_bin].to(torch.int32)
    return _from_dev(ig, device_mode), _from_dev(ifg, device_mode), _from_dev(cnt, device_mode)


# ==================================================================================================
# a8  rc.GroupByAllPack32 (riptable/rt_numpy.py:1891; riptable/rt_grouping.py:3443-3454)
def GroupByAllPack32(values, ikey, iGroup, iFirstGroup, nCountGroup, unique_rows, funcList, binLowList, binHighList,
                     func_param=0):
    values = _as_list(values)
    device_mode = any(_is_device(x) for x in (ikey, iGroup, iFirstGroup, nCountGroup)) or any(_is_device(v) for v in values)
    ig, ifg, cnt = (_to_dev(x) for x in (iGroup, iFirstGroup, nCountGroup))
    for name, d in (("iGroup", ig), ("iFirstGroup", ifg), ("nCountGroup", cnt)):
        if d.dtype != np.dtype(np.int32):
            raise TypeError(f"GroupByAllPack32: {name} must be int32")
    nb = int(unique_rows) + 1
    if ifg.n < nb or cnt.n < nb:
        raise ValueError("GroupByAllPack32: iFirstGroup / nCountGroup shorter than unique_rows + 1")
    out = []
    kdev = None
    for v, f, lo, hi in zip(values, funcList, binLowList, binHighList):
        f = int(f)
        if f == GB_FUNCTIONS.GB_ROLLING_QUANTILE:
            # 8f.2: func_param = int(q * 1e9) + window * (1e9 + 1)  (riptable/rt_utils.py:1172-1184)
            if kdev is None:
                kdev = _ikey_dev(ikey)
            d = _to_dev(v)
            if d.dtype.kind not in "biuf" or d.dtype == np.float16:
                out.append(None)
                continue
            fp = int(func_param)
            window, q = fp // 1_000_000_001, (fp % 1_000_000_001) / 1e9
            res = DevArray.empty(kdev.n, np.float64)
            check(lib.rtb_groupby_rolling_quantile(C.c_void_p(d.ptr), dtype_code(d.dtype), C.c_void_p(kdev.ptr), dtype_code(kdev.dtype),
                                                   C.c_void_p(ig.ptr), C.c_void_p(ifg.ptr), C.c_void_p(cnt.ptr), kdev.n,
                                                   int(unique_rows), int(window), float(q), C.c_void_p(res.ptr), _stream()))
            out.append(_from_dev(res, device_mode))
            continue
        if GB_FUNCTIONS.GB_ROLLING_SUM <= f <= GB_FUNCTIONS.GB_ROLLING_NANMEAN:
            # 8f.2: row-shaped rolling results; func_param is the window / period (riptable/rt_groupbyops.py:2941-3153)
            if kdev is None:
                kdev = _ikey_dev(ikey)
            if f == GB_FUNCTIONS.GB_ROLLING_COUNT:
                d, vcode, isz = None, RTB_BYTES, 0
                ocode = lib.rtb_rolling_out_dtype(f, 0)
            else:
                d = _to_dev(v)
                vcode = dtype_code(d.dtype) if d.dtype != np.float16 else -1
                isz = d.dtype.itemsize
                ocode = lib.rtb_rolling_out_dtype(f, vcode) if vcode >= 0 else -1
            if ocode < 0:
                out.append(None)
                continue
            odt = d.dtype if f == GB_FUNCTIONS.GB_ROLLING_SHIFT else code_dtype(ocode)
            res = DevArray.empty(kdev.n, odt)
            check(lib.rtb_groupby_rolling(f, C.c_void_p(d.ptr) if d is not None else None, vcode, isz, C.c_void_p(kdev.ptr),
                                          dtype_code(kdev.dtype), C.c_void_p(ig.ptr), C.c_void_p(ifg.ptr), C.c_void_p(cnt.ptr),
                                          kdev.n, int(unique_rows), int(func_param), C.c_void_p(res.ptr), _stream()))
            out.append(_from_dev(res, device_mode))
            continue
        if f == GB_FUNCTIONS.GB_MODE:
            # 8f.2: most frequent valid value per bin, in the value's own dtype (riptable/rt_groupbyops.py:1297-1363)
            if kdev is None:
                kdev = _ikey_dev(ikey)
            d = _to_dev(v)
            if d.dtype.kind not in "biuf" or d.dtype == np.float16:
                out.append(None)
                continue
            res = DevArray.empty(nb, d.dtype)
            ws = _ws(lib.rtb_quantile_workspace_bytes(kdev.n, int(unique_rows)))
            check(lib.rtb_groupby_mode(C.c_void_p(d.ptr), dtype_code(d.dtype), C.c_void_p(kdev.ptr), dtype_code(kdev.dtype),
                                       C.c_void_p(ifg.ptr), C.c_void_p(cnt.ptr), kdev.n, int(unique_rows), int(lo), int(hi),
                                       C.c_void_p(res.ptr), C.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
            out.append(_from_dev(res, device_mode))
            continue
        if f in (GB_FUNCTIONS.GB_MEDIAN, GB_FUNCTIONS.GB_QUANTILE_MULT):
            # 8f.2: median "auto handles nan"; quantile decodes riptable's integer parameter
            # (riptable/rt_groupbyops.py:1717-1745: int(q * 1e9) + is_nan * (1e9 + 1))
            if kdev is None:
                kdev = _ikey_dev(ikey)
            d = _to_dev(v)
            if d.dtype.kind not in "biuf" or d.dtype == np.float16:
                out.append(None)
                continue
            if f == GB_FUNCTIONS.GB_MEDIAN:
                q, nan_aware = 0.5, 1
            else:
                fp = int(func_param)
                nan_aware = 1 if fp >= 1_000_000_001 else 0
                q = (fp - nan_aware * 1_000_000_001) / 1e9
            res = DevArray.empty(nb, np.float64)
            ws = _ws(lib.rtb_quantile_workspace_bytes(kdev.n, int(unique_rows)))
            check(lib.rtb_groupby_quantile(C.c_void_p(d.ptr), dtype_code(d.dtype), C.c_void_p(kdev.ptr), dtype_code(kdev.dtype),
                                           C.c_void_p(ifg.ptr), C.c_void_p(cnt.ptr), kdev.n, int(unique_rows), float(q), nan_aware,
                                           int(lo), int(hi), C.c_void_p(res.ptr), C.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
            out.append(_from_dev(res, device_mode))
            continue
        if f not in (GB_FUNCTIONS.GB_FIRST, GB_FUNCTIONS.GB_NTH, GB_FUNCTIONS.GB_LAST):
            raise NotImplementedError(f"GroupByAllPack32: function {f} is outside this build's hot path")
        d = _to_dev(v)
        res = DevArray.empty(nb, d.dtype)
        nth = int(func_param) if f == GB_FUNCTIONS.GB_NTH else 0
        check(lib.rtb_groupby_pick(C.c_void_p(d.ptr), dtype_code(d.dtype), d.dtype.itemsize, C.c_void_p(ig.ptr),
                                   C.c_void_p(ifg.ptr), C.c_void_p(cnt.ptr), int(unique_rows), f, nth, int(lo), int(hi),
                                   C.c_void_p(res.ptr), _stream()))
        out.append(_from_dev(res, device_mode))
    return tuple(out)


# ==================================================================================================
# a9 / 8f.2  rc.EmaAll32 (riptable/rt_grouping.py:3430): GB_CUMSUM, GB_CUMPROD, GB_CUM[NAN]MIN, GB_CUM[NAN]MAX
_CUM_FUNCS = (GB_FUNCTIONS.GB_CUMSUM, GB_FUNCTIONS.GB_CUMPROD, GB_FUNCTIONS.GB_CUMNANMAX, GB_FUNCTIONS.GB_CUMNANMIN,
              GB_FUNCTIONS.GB_CUMMAX, GB_FUNCTIONS.GB_CUMMIN)


_EMA_FUNCS = (GB_FUNCTIONS.GB_EMADECAY, GB_FUNCTIONS.GB_EMANORMAL, GB_FUNCTIONS.GB_EMAWEIGHTED)


def EmaAll32(values, ikey, unique_rows, funcNum, func_param):
    func = int(funcNum)
    if func not in _CUM_FUNCS and func not in _EMA_FUNCS:
        raise NotImplementedError(f"EmaAll32: function {func} is outside this build's hot path")
    values = _as_list(values)
    rate, time, filt, reset = 0.0, None, None, None
    if isinstance(func_param, (tuple, list)) and len(func_param) >= 4:
        rate, time, filt, reset = func_param[0], func_param[1], func_param[2], func_param[3]
    device_mode = _is_device(ikey) or any(_is_device(v) for v in values)
    k = _ikey_dev(ikey)
    fdev = _bool_dev(filt, k.n) if filt is not None else None
    rdev = _bool_dev(reset, k.n, "reset_filter") if reset is not None else None
    tdev = None
    if func in _EMA_FUNCS:
        # 8f.2: riptable/rt_groupbyops.py:3279-3498 — func_param = (decay_rate, time, filter, reset_filter)
        if rate is None:
            raise ValueError("EmaAll32: the ema functions need a decay_rate")
        if func != GB_FUNCTIONS.GB_EMAWEIGHTED:
            if time is None:
                raise ValueError("EmaAll32: ema_decay / ema_normal need a time array")
            tdev = _to_dev(time)
            if tdev.n != k.n:
                raise ValueError(f"EmaAll32: the time array length {tdev.n} must match the key length {k.n}")
            if tdev.dtype.kind not in "iuf" or tdev.dtype == np.float16:
                raise TypeError("EmaAll32: the time array must be an integer or floating point array")
    out = []
    for v in values:
        d = _to_dev(v)
        vcode = dtype_code(d.dtype) if d.dtype.kind in "biuf" and d.dtype != np.float16 else RTB_BYTES
        if func in _EMA_FUNCS:
            ocode = lib.rtb_ema_out_dtype(func, vcode) if vcode < RTB_BYTES else -1
        else:
            ocode = lib.rtb_cumop_out_dtype(func, vcode) if vcode < RTB_BYTES else -1
        if ocode < 0:
            out.append(None)
            continue
        if d.n != k.n:
            raise ValueError("EmaAll32: value and key lengths differ")
        res = DevArray.empty(k.n, code_dtype(ocode))
        fptr = C.c_void_p(fdev.ptr) if fdev is not None else None
        rptr = C.c_void_p(rdev.ptr) if rdev is not None else None
        if func in _EMA_FUNCS:
            ws = _ws(lib.rtb_ema_workspace_bytes(k.n, int(unique_rows)))
            check(lib.rtb_groupby_ema(func, C.c_void_p(d.ptr), vcode, C.c_void_p(k.ptr), dtype_code(k.dtype), k.n, int(unique_rows),
                                      C.c_void_p(tdev.ptr) if tdev is not None else None,
                                      dtype_code(tdev.dtype) if tdev is not None else 0, float(rate), fptr, rptr,
                                      C.c_void_p(res.ptr), C.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
        else:
            ws = _ws(lib.rtb_cumop_workspace_bytes(func, k.n, int(unique_rows), int(rdev is not None)))
            check(lib.rtb_groupby_cumop(func, C.c_void_p(d.ptr), vcode, C.c_void_p(k.ptr), dtype_code(k.dtype), k.n,
                                        int(unique_rows), fptr, rptr,
                                        C.c_void_p(res.ptr), C.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
        out.append(_from_dev(res, device_mode))
    return out


# ==================================================================================================
# a10  rc.MakeiFirst / rc.MakeiNext (riptable/rt_numpy.py:1791-1854)
def MakeiFirst(key, n, filter=None, mode=0):
    device_mode = _is_device(key)
    k = _ikey_dev(key)
    f = _bool_dev(filter, k.n) if filter is not None else None
    out = DevArray.empty(int(n) + 1, np.int32)
    check(lib.rtb_make_ifirst(C.c_void_p(k.ptr), dtype_code(k.dtype), k.n, int(n), C.c_void_p(f.ptr) if f is not None else None,
                              int(mode), C.c_void_p(out.ptr), _stream()))
    return _from_dev(out, device_mode)


def MakeiNext(key, n, mode=0):
    device_mode = _is_device(key)
    k = _ikey_dev(key)
    out = DevArray.empty(k.n, np.int32)
    ws = _ws(lib.rtb_make_inext_workspace_bytes(k.n, int(n)))
    check(lib.rtb_make_inext(C.c_void_p(k.ptr), dtype_code(k.dtype), k.n, int(n), int(mode), C.c_void_p(out.ptr),
                             C.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
    return _from_dev(out, device_mode)


# ==================================================================================================
# a11  rc.LexSort32 / rc.LexSort64 (riptable/rt_numpy.py:2670-2671, 2200-2202)
def LexSort32(list_arrays, cutoffs=None, index=None, ascending=True):
    """Stable lexicographic argsort; the LAST array is the primary key (numpy.lexsort convention), NaN last.
    cutoffs: every partition is sorted as an array of its own, out[c[p-1]:c[p]] = its partition-relative order
    (unpinned by the reference's tests; derived from the pinned hash case, see oracle.LexSort32)."""
    list_arrays = _as_list(list_arrays)
    if not isinstance(list_arrays, list) or len(list_arrays) == 0:
        raise ValueError("lexsort needs a list of arrays")
    device_mode = any(_is_device(a) for a in list_arrays) or _is_device(index)
    cols = [_to_dev(a) for a in list_arrays]
    n = cols[0].n
    if any(c.n != n for c in cols):
        raise ValueError("lexsort: all keys must have the same length")
    if n > 2_000_000_000:
        raise ValueError("LexSort32: more than 2e9 rows; call LexSort64 (int64 row numbers)")
    if isinstance(ascending, (bool, np.bool_)):
        asc = np.full(len(cols), 1 if ascending else 0, dtype=np.int8)
    else:
        asc = np.asarray([1 if a else 0 for a in ascending], dtype=np.int8)
        if len(asc) != len(cols):
            raise ValueError("ascending must have one entry per key")
    cut = None
    if cutoffs is not None:
        if index is not None:
            raise ValueError("lexsort: cutoffs and index (a filter) cannot be combined")
        if n == 0:
            return _from_dev(DevArray.empty(0, np.int32), device_mode)
        # the partition id is one more key, more significant than all others: one sort, partitions end to end
        cut, cdev, pid = _cutoffs_dev(cutoffs, n)
        cols = cols + [_to_dev(pid)]
        asc = np.concatenate([asc, np.ones(1, dtype=np.int8)])
    idx = None
    m = n
    if index is not None:
        idx = _to_dev(index)
        if idx.dtype != np.dtype(np.int32):
            idx = _to_dev(idx.numpy().astype(np.int32))
        m = idx.n
    out = DevArray.empty(m, np.int32)
    if m == 0:
        return _from_dev(out, device_mode)
    isz = (C.c_int32 * len(cols))(*[c.dtype.itemsize for c in cols])
    ws = _ws(lib.rtb_lexsort_workspace_bytes(n, m, len(cols), isz))
    check(lib.rtb_lexsort32(_cols(cols), len(cols), n, C.c_void_p(idx.ptr) if idx is not None else None, m,
                            asc.ctypes.data_as(C.POINTER(C.c_int8)), C.c_void_p(out.ptr), C.c_void_p(ws.data_ptr()),
                            ws.numel(), _stream()))
    if cut is not None:  # global rows -> partition-relative rows (position i lies in partition pid[i])
        check(lib.rtb_partition_shift(C.c_void_p(out.ptr), n, C.c_void_p(pid.data_ptr()), C.c_void_p(cdev.data_ptr()), len(cut), 1, _stream()))
    return _from_dev(out, device_mode)


def LexPartition(list_arrays, splitter_arrays, splitter_rows, row_offset=0, ascending=True):
    """Sample-sort helper (device mode): for every row, how many of the splitter rows sort at or before it under
    LexSort32's order with the global row id as the final tie-break.  -> int32 CUDA tensor."""
    cols = [_to_dev(a) for a in _as_list(list_arrays)]
    spl = [_to_dev(a) for a in _as_list(splitter_arrays)]
    n = cols[0].n
    nsplit = spl[0].n if spl else 0
    if isinstance(ascending, (bool, np.bool_)):
        asc = np.full(len(cols), 1 if ascending else 0, dtype=np.int8)
    else:
        asc = np.asarray([1 if a else 0 for a in ascending], dtype=np.int8)
    srows = torch.as_tensor(splitter_rows, dtype=torch.int64, device=_device()).contiguous()
    out = torch.empty(max(n, 4), dtype=torch.int32, device=_device())
    check(lib.rtb_lex_partition(_cols(cols), len(cols), n, asc.ctypes.data_as(C.POINTER(C.c_int8)),
                                _cols(spl) if nsplit else None, nsplit, C.c_void_p(srows.data_ptr()) if nsplit else None,
                                int(row_offset), C.c_void_p(out.data_ptr()), _stream()))
    return out[:n]


def LexSort64(list_arrays, cutoffs=None, index=None, ascending=True):
    """rc.LexSort64 (riptable/rt_numpy.py:2199-2200, 2669-2670: what riptable calls above 2e9 rows): the same stable
    lexicographic argsort with an int64 result (and an int64 `index`), for up to 2^32 - 2 rows."""
    list_arrays = _as_list(list_arrays)
    if not isinstance(list_arrays, list) or len(list_arrays) == 0:
        raise ValueError("lexsort needs a list of arrays")
    n0 = len(list_arrays[0])
    if n0 <= 2_000_000_000 and (index is None or len(index) <= 2_000_000_000) and cutoffs is not None:
        # partitions: the 32-bit path does the partition bookkeeping; only the result's dtype differs
        out = LexSort32(list_arrays, cutoffs=cutoffs, index=index, ascending=ascending)
        return out.to(torch.int64) if isinstance(out, torch.Tensor) else out.astype(np.int64)
    if cutoffs is not None:
        raise ValueError("LexSort64: cutoffs are provided up to 2e9 rows")
    device_mode = any(_is_device(a) for a in list_arrays) or _is_device(index)
    cols = [_to_dev(a) for a in list_arrays]
    n = cols[0].n
    if any(c.n != n for c in cols):
        raise ValueError("lexsort: all keys must have the same length")
    if n > 4_294_967_294:
        raise ValueError("LexSort64: at most 2^32 - 2 rows")
    if isinstance(ascending, (bool, np.bool_)):
        asc = np.full(len(cols), 1 if ascending else 0, dtype=np.int8)
    else:
        asc = np.asarray([1 if a else 0 for a in ascending], dtype=np.int8)
        if len(asc) != len(cols):
            raise ValueError("ascending must have one entry per key")
    idx, m = None, n
    if index is not None:
        idx = _to_dev(index)
        if idx.dtype != np.dtype(np.int64):
            idx = _to_dev(idx.torch().to(torch.int64))
        m = idx.n
    out = DevArray.empty(m, np.int64)
    if m == 0:
        return _from_dev(out, device_mode)
    isz = (C.c_int32 * len(cols))(*[c.dtype.itemsize for c in cols])
    ws = _ws(lib.rtb_lexsort64_workspace_bytes(n, m, len(cols), isz))
    check(lib.rtb_lexsort64(_cols(cols), len(cols), n, C.c_void_p(idx.ptr) if idx is not None else None, m,
                            asc.ctypes.data_as(C.POINTER(C.c_int8)), C.c_void_p(out.ptr), C.c_void_p(ws.data_ptr()),
                            ws.numel(), _stream()))
    return _from_dev(out, device_mode)


# ==================================================================================================
# a12  rc.GroupFromLexSort (riptable/rt_numpy.py:2207)
def GroupFromLexSort(iGroup, value_array, cutoffs=None, base_index=1):
    """-> (iKey int32[N], iFirstKey int32[U], nCountGroup int32[U+1] with slot 0 = 0).
    cutoffs: -> (iKey, iFirstKey, nCountGroup, nUniqueCutoffs): per-partition results laid end to end -- bin numbering
    restarts in every partition, iFirstKey holds partition-relative rows, nUniqueCutoffs the cumulative unique counts
    (unpinned by the reference's tests; derived from the pinned hash case, see oracle.GroupFromLexSort)."""
    if isinstance(value_array, np.ndarray) and value_array.dtype.names:
        value_array = [np.ascontiguousarray(value_array[nm]) for nm in value_array.dtype.names]
    vals = _as_list(value_array)
    device_mode = _is_device(iGroup) or any(_is_device(v) for v in vals)
    ig = _to_dev(iGroup)
    if ig.dtype != np.dtype(np.int32):
        raise TypeError("GroupFromLexSort: iGroup must be int32")
    cols = [_to_dev(v) for v in vals]
    n = cols[0].n
    m = ig.n
    cut = None
    base = int(base_index)
    if cutoffs is not None:
        if m != n:
            raise ValueError("GroupFromLexSort: cutoffs need an iGroup that covers every row (no filter)")
        cut, cdev, pid = _cutoffs_dev(cutoffs, n)
        if n:
            # partition-relative sorted rows -> global rows; the partition id becomes one more key column, so bins never
            # run across a partition boundary
            g = ig.torch().clone()
            check(lib.rtb_partition_shift(C.c_void_p(g.data_ptr()), n, C.c_void_p(pid.data_ptr()), C.c_void_p(cdev.data_ptr()), len(cut), 0, _stream()))
            ig = _to_dev(g)
            cols = cols + [_to_dev(pid)]
        base = 1
    ikey = DevArray.empty(n, np.int32)
    ifk = DevArray.empty(max(m, 1), np.int32)
    cnt = DevArray.empty(m + 1, np.int32)
    u = C.c_int64(0)
    ws = _ws(lib.rtb_group_from_lexsort_workspace_bytes(n, m))
    check(lib.rtb_group_from_lexsort(C.c_void_p(ig.ptr), m, _cols(cols), len(cols), n, base, C.c_void_p(ikey.ptr),
                                     C.c_void_p(ifk.ptr), C.c_void_p(cnt.ptr), C.byref(u), C.c_void_p(ws.data_ptr()),
                                     ws.numel(), _stream()))
    U = int(u.value)
    ifk = DevArray(ifk.buf, np.int32, U)
    cnt = DevArray(cnt.buf, np.int32, U + 1)
    if cut is None:
        return _from_dev(ikey, device_mode), _from_dev(ifk, device_mode), _from_dev(cnt, device_mode)
    P = len(cut)
    ucount = torch.empty(P, dtype=torch.int32, device=_device())
    check(lib.rtb_partition_count_rows(C.c_void_p(ifk.ptr), U, C.c_void_p(pid.data_ptr()), P, C.c_void_p(ucount.data_ptr()), _stream()))
    ucut = torch.cumsum(ucount.to(torch.int64), 0)
    bin_base = (ucut - ucount).to(torch.int32).contiguous()
    check(lib.rtb_partition_localize(C.c_void_p(ikey.ptr), n, C.c_void_p(ifk.ptr), U, C.c_void_p(pid.data_ptr()),
                                     C.c_void_p(bin_base.data_ptr()), C.c_void_p(cdev.data_ptr()), P, int(base_index), _stream()))
    return (_from_dev(ikey, device_mode), _from_dev(ifk, device_mode), _from_dev(cnt, device_mode),
            ucut if device_mode else ucut.cpu().numpy())


# ==================================================================================================
# a13  rc.ReverseShuffle (riptable/rt_grouping.py:1659)
def ReverseShuffle(perm):
    device_mode = _is_device(perm)
    p = _to_dev(perm)
    if p.dtype != np.dtype(np.int32):
        if p.dtype.kind != "i":
            raise TypeError("ReverseShuffle: permutation must be an integer array")
        p = _to_dev(p.numpy().astype(np.int32))
    out = DevArray.empty(p.n, np.int32)
    check(lib.rtb_reverse_shuffle(C.c_void_p(p.ptr), p.n, C.c_void_p(out.ptr), _stream()))
    return _from_dev(out, device_mode)


# 8f.3  rc.ReIndexGroups (riptable/rt_grouping.py:378; semantics: the Python fallback at :380-407)
def ReIndexGroups(ikey, uikey, u_cutoffs, i_cutoffs):
    device_mode = _is_device(ikey) or _is_device(uikey)
    k = _to_dev(ikey)
    u = _to_dev(uikey)
    if k.dtype.kind != "i" or u.dtype.kind != "i":
        raise TypeError("ReIndexGroups: ikey and uikey must be signed integer arrays")
    kdt = k.dtype
    if k.dtype != np.dtype(np.int32):
        k = _to_dev(k.numpy().astype(np.int32))
    if u.dtype != np.dtype(np.int32):
        u = _to_dev(u.numpy().astype(np.int32))
    uc = np.ascontiguousarray(np.asarray(u_cutoffs, dtype=np.int64))
    ic = np.ascontiguousarray(np.asarray(i_cutoffs, dtype=np.int64))
    if uc.ndim != 1 or uc.shape != ic.shape:
        raise ValueError("ReIndexGroups: u_cutoffs and i_cutoffs must be 1-d arrays of the same length")
    out = DevArray.empty(k.n, np.int32)
    check(lib.rtb_reindex_groups(C.c_void_p(k.ptr), k.n, C.c_void_p(u.ptr), u.n, uc.ctypes.data_as(C.POINTER(C.c_int64)),
                                 ic.ctypes.data_as(C.POINTER(C.c_int64)), len(uc), C.c_void_p(out.ptr), _stream()))
    res = _from_dev(out, device_mode)
    return res if device_mode or kdt == np.dtype(np.int32) else res.astype(kdt)


__all__ = ["MultiKeyGroupBySum32", "GroupByStream", "ReIndexGroups", "LexPartition", "MultiKeyGroupBy32", "CombineFilter", "CombineAccum1Filter", "CombineAccum2Filter", "GroupByAll32", "BinCount",
           "GroupFromBinCount", "GroupByPack32", "GroupByAllPack32", "EmaAll32", "MakeiFirst", "MakeiNext", "LexSort32",
           "LexSort64", "GroupFromLexSort", "ReverseShuffle", "IsMember32", "IsMemberCategorical", "IsMemberCategoricalFixup", "MultiKeyIsMember32", "LedgerFunction"]


# ==================================================================================================
# 8f.1  rc.IsMember32 / rc.MultiKeyIsMember32 (riptable/rt_numpy.py:1186-1393)
def _concat_dev(b: DevArray, a: DevArray) -> DevArray:
    """[b ; a] as one device column; fixed-width strings are padded to the wider item first."""
    if b.dtype.kind in "SU" or a.dtype.kind in "SU":
        if a.dtype.kind != b.dtype.kind:
            raise TypeError("ismember: string keys must be matched with string keys of the same kind")
        w = max(a.dtype.itemsize, b.dtype.itemsize)
        dt = np.dtype(f"{a.dtype.kind}{w // (4 if a.dtype.kind == 'U' else 1)}")
        out = torch.zeros(max((a.n + b.n) * w, 16), dtype=torch.uint8, device=_device())
        for arr, lo in ((b, 0), (a, b.n)):
            if arr.n:
                isz = arr.dtype.itemsize
                out[lo * w: (lo + arr.n) * w].view(arr.n, w)[:, :isz] = arr.buf[: arr.n * isz].view(arr.n, isz)
        return DevArray(out, dt, a.n + b.n)
    if a.dtype != b.dtype:
        raise TypeError("ismember: numeric keys must share a dtype (riptable casts before calling rc)")
    isz = a.dtype.itemsize
    out = torch.empty(max((a.n + b.n) * isz, 16), dtype=torch.uint8, device=_device())
    out[: b.n * isz] = b.buf[: b.n * isz]
    out[b.n * isz: (a.n + b.n) * isz] = a.buf[: a.n * isz]
    return DevArray(out, a.dtype, a.n + b.n)


def _ismember(a_cols, b_cols, base_index=0):
    """Hash-group the rows of b followed by the rows of a: a row of `a` is a member iff its bin's first row lies in
    the b part, and that first row is the index of its first occurrence in b.  NaN keys never match."""
    from .enums import int_dtype_from_len, invalid_for

    device_mode = any(_is_device(c) for c in list(a_cols) + list(b_cols))
    if len(a_cols) != len(b_cols) or not a_cols:
        raise ValueError("ismember: a and b must have the same number of key columns")
    acs, bcs = [_to_dev(c) for c in a_cols], [_to_dev(c) for c in b_cols]
    na, nb = acs[0].n, bcs[0].n
    idt = int_dtype_from_len(nb + base_index)
    tdt = _TORCH_OF[idt]
    inv = 0 if base_index else int(invalid_for(idt))  # base_index 1: positions count from 1 and 0 means absent
    dev = _device()
    if na == 0 or nb == 0:
        found = torch.zeros(na, dtype=torch.bool, device=dev)
        idx = torch.full((na,), inv, dtype=tdt, device=dev)
    elif (len(acs) == 1 and acs[0].dtype == bcs[0].dtype and acs[0].dtype.kind in "biufmM" and acs[0].dtype.itemsize in (1, 2, 4, 8)
          and acs[0].dtype != np.float16):
        # one numeric key column: build the table from b, look the rows of a up (no concatenation, no inserts)
        found_u8 = torch.empty(max(na, 1), dtype=torch.uint8, device=dev)[:na]
        idx = torch.empty(max(na, 1), dtype=tdt, device=dev)[:na]
        ws = _ws(lib.rtb_ismember_workspace_bytes(nb))
        check(lib.rtb_ismember32(C.c_void_p(acs[0].ptr), na, C.c_void_p(bcs[0].ptr), nb, dtype_code(acs[0].dtype), int(base_index),
                                 dtype_code(idt), C.c_void_p(found_u8.data_ptr()), C.c_void_p(idx.data_ptr()),
                                 C.c_void_p(ws.data_ptr()), ws.numel(), _stream()))
        found = found_u8.view(torch.bool)
    else:
        cat = [_concat_dev(bc, ac) for ac, bc in zip(acs, bcs)]
        valid = None
        for c in cat:
            if c.dtype.kind == "f" and c.dtype != np.float16:  # NaN never matches: those rows stay out of the grouping
                if valid is None:
                    valid = torch.empty(max(c.n, 4), dtype=torch.uint8, device=dev)[: c.n]
                    mode = 0
                else:
                    mode = 1
                check(lib.rtb_valid_mask(C.c_void_p(c.ptr), dtype_code(c.dtype), c.n, mode, C.c_void_p(valid.data_ptr()), _stream()))
        ik, ifk, u = MultiKeyGroupBy32(cat, 0, None if valid is None else valid.view(torch.bool), 2)
        found_u8 = torch.empty(max(na, 1), dtype=torch.uint8, device=dev)[:na]
        idx = torch.empty(max(na, 1), dtype=tdt, device=dev)[:na]
        ika = ik[nb:].contiguous()
        if ika.data_ptr() % 4:
            ika = ika.clone()
        check(lib.rtb_ismember_finish(C.c_void_p(ika.data_ptr()), na, C.c_void_p(ifk.data_ptr()) if u else None, int(u), nb,
                                      int(base_index), dtype_code(idt), C.c_void_p(found_u8.data_ptr()), C.c_void_p(idx.data_ptr()),
                                      _stream()))
        found = found_u8.view(torch.bool)
    if device_mode:
        return found, idx
    return found.cpu().numpy(), idx.cpu().numpy()


def IsMember32(a, b, h=2, hint_size=0):
    """-> (bool[len(a)], index of the first occurrence of a[i] in b; the index dtype (by len(b)) 's invalid where absent)."""
    return _ismember([a], [b])


def IsMemberCategorical(a, b, h=2, hint_size=0):
    """rc.IsMemberCategorical (riptable/rt_numpy.py:1388-1389, `ismember(..., base_index=1)`): the position is 1-based
    and 0 where absent, ready to be used as a Categorical's codes (riptable/rt_grouping.py:959-963)."""
    return _ismember([a], [b], 1)


def IsMemberCategoricalFixup(a_codes, b_codes, on_unique, num_unique_b, a_base_index=1, b_base_index=1):
    """rc.IsMemberCategoricalFixup (riptable/rt_numpy.py:1302-1305): ismember of two Categoricals from the match of their
    uniques.  a_codes / b_codes: the Categoricals' underlying integer arrays; on_unique: int32 position of each unique of
    `a` in the uniques of `b` (invalid / negative where absent).  -> (bool[len(a)], int32 first row of b holding the same
    category, the int32 invalid where there is none) == ismember(a.expand_array, b.expand_array)
    (riptable/tests/test_ismember.py:162-195)."""
    device_mode = _is_device(a_codes) or _is_device(b_codes)
    a, b = _ikey_dev(a_codes), _ikey_dev(b_codes)
    ou = _to_dev(on_unique)
    if ou.dtype != np.dtype(np.int32):
        ou = _to_dev((ou.torch() if ou.dtype in _TORCH_OF else torch.from_numpy(ou.numpy())).to(torch.int32))
    nub = int(num_unique_b)
    dev = _device()
    # first row of b per CODE of b (codes 0 .. nub + b_base - 1); codes outside that range are ignored
    nslots = nub + int(b_base_index)
    b_first = DevArray.empty(nslots + 1, np.int32)
    check(lib.rtb_first_row_per_code(C.c_void_p(b.ptr), dtype_code(b.dtype), b.n, nslots, C.c_void_p(b_first.ptr), _stream()))
    found = torch.empty(max(a.n, 1), dtype=torch.uint8, device=dev)[: a.n]
    idx = torch.empty(max(a.n, 1), dtype=torch.int32, device=dev)[: a.n]
    check(lib.rtb_ismember_categorical_fixup(C.c_void_p(a.ptr), dtype_code(a.dtype), a.n, C.c_void_p(ou.ptr), ou.n,
                                             C.c_void_p(b_first.ptr), nub, int(a_base_index), int(b_base_index),
                                             C.c_void_p(found.data_ptr()), C.c_void_p(idx.data_ptr()), _stream()))
    found = found.view(torch.bool)
    if device_mode:
        return found, idx
    return found.cpu().numpy(), idx.cpu().numpy()


def MultiKeyIsMember32(a_tuple, b_tuple, hint_size=0):
    """rc.MultiKeyIsMember32((a_cols,), (b_cols,), hint_size) (riptable/rt_numpy.py:1325)."""
    return _ismember(list(a_tuple[0]), list(b_tuple[0]))


def LedgerFunction(func, *args, **kwargs):
    """rc.LedgerFunction(f, *a, **k) == f(*a, **k) with timing recorded (riptable/rt_numpy.py:153,1871)."""
    return func(*args, **kwargs)
