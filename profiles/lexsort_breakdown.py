#!/usr/bin/env python
"""Where the time of the C4 lexsort goes: library profile slots (CUDA-event totals per kernel class)."""
import ctypes as C
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402
from riptable_b200 import _lib  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(12345)
a = torch.randint(0, 1000, (n,), device="cuda", generator=g, dtype=torch.int64)
f = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
c = torch.randint(0, 1_000_000, (n,), device="cuda", generator=g, dtype=torch.int32)
rc.LexSort32([c, f, a]); torch.cuda.synchronize()
lib = _lib.lib
lib.rtb_profile_enable(1); lib.rtb_profile_reset()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); p = rc.LexSort32([c, f, a]); e1.record(); torch.cuda.synchronize()
lib.rtb_profile_enable(0)
print("total ms %.2f" % e0.elapsed_time(e1))
names = {2: "radix_hist", 3: "radix_scatter", 4: "segscan", 5: "gather (sort codes by the current order)"}
for i, nm in names.items():
    ms, la, un = C.c_double(), C.c_uint64(), C.c_uint64()
    if lib.rtb_profile_get(i, C.byref(ms), C.byref(la), C.byref(un)) == 0 and la.value:
        print("%-45s %8.2f ms  %3d launches" % (nm, ms.value, la.value))
