#!/usr/bin/env python
"""Small-input call latency: the C entry point alone (preallocated buffers) vs the Python shim around it."""
import ctypes as C
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402
from riptable_b200 import _lib  # noqa: E402
from riptable_b200._lib import lib, rtb_column  # noqa: E402

for n in (1000, 1_000_000, 10_000_000):
    k = torch.randint(0, 1000, (n,), device="cuda", dtype=torch.int64)
    mu = max(1 << 20, n // 2)
    mu = min(mu, n)
    wsb = lib.rtb_groupby_workspace_bytes_rows(n, mu, 8, 0)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    ik = torch.empty(n + 4, dtype=torch.int32, device="cuda")
    ifk = torch.empty(mu + 4, dtype=torch.int32, device="cuda")
    col = (rtb_column * 1)()
    col[0] = rtb_column(C.c_void_p(k.data_ptr()), 7, 8)
    u, need = C.c_int64(), C.c_int64()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call():
        return lib.rtb_multikey_groupby32(col, 1, n, 0, None, None, 0, C.c_void_p(ik.data_ptr()), C.c_void_p(ifk.data_ptr()), mu,
                                          C.byref(u), None, C.byref(need), C.c_void_p(ws.data_ptr()), wsb, stream)

    for _ in range(5):
        assert call() == 0
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(50):
        call()
    torch.cuda.synchronize()
    tc = (time.perf_counter() - t) / 50
    for _ in range(5):
        rc.MultiKeyGroupBy32([k], 0, None, 2)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(50):
        rc.MultiKeyGroupBy32([k], 0, None, 2)
    torch.cuda.synchronize()
    tp = (time.perf_counter() - t) / 50
    print(f"rows {n:>9}: C entry point {tc * 1e6:6.0f} us, rc.MultiKeyGroupBy32 {tp * 1e6:6.0f} us, uniques {u.value}")
