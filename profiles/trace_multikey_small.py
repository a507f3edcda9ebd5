#!/usr/bin/env python
"""Two small integer key columns (int32, int32): the packed-row route, per-round trace."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RTB_GB_TRACE"] = "1"
import riptable_b200.rc as rc  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
ua = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ub = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
g = torch.Generator(device="cuda"); g.manual_seed(1)
adt = torch.int64 if os.environ.get("A64") else torch.int32
a = torch.randint(0, ua, (n,), device="cuda", generator=g, dtype=adt)
b = torch.randint(0, ub, (n,), device="cuda", generator=g, dtype=torch.int32)
rc.MultiKeyGroupBy32([a, b], 0, None, 2); torch.cuda.synchronize()
print("---- second call", file=sys.stderr, flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ik, ifk, u = rc.MultiKeyGroupBy32([a, b], 0, None, 2); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(("(int64, int32)" if os.environ.get("A64") else "(int32, int32)") + ", %d rows, %d uniques: %.2f ms = %.1f G rows/s" % (n, u, ms, n / ms / 1e6), file=sys.stderr)
