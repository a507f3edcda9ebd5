#!/usr/bin/env python
"""Per-round trace (RTB_GB_TRACE=1) of the high-cardinality hash grouping: C5 (int64 key, N/10 uniques)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RTB_GB_TRACE"] = "1"
import riptable_b200.rc as rc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda")
g.manual_seed(12345)
u = int(sys.argv[2]) if len(sys.argv) > 2 else n // 10
uv = torch.randint(-(2**62), 2**62, (u,), dtype=torch.int64, device="cuda", generator=g)
k = torch.empty(n, dtype=torch.int64, device="cuda")
for s in range(0, n, 1 << 27):
    e = min(n, s + (1 << 27))
    torch.index_select(uv, 0, torch.randint(0, u, (e - s,), device="cuda", generator=g), out=k[s:e])
rc.MultiKeyGroupBy32([k], 0, None, 2)
torch.cuda.synchronize()
print("---- second call", file=sys.stderr, flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ik, ifk, U = rc.MultiKeyGroupBy32([k], 0, None, 2)
e1.record()
torch.cuda.synchronize()
print("total ms", e0.elapsed_time(e1), "uniques", U, file=sys.stderr)
