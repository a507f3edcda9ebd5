#!/usr/bin/env python
"""RTB_GB_TRACE of low-cardinality int64 grouping (100 uniques, 100 M rows)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RTB_GB_TRACE"] = "1"
import riptable_b200.rc as rc  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
u = int(sys.argv[2]) if len(sys.argv) > 2 else 100
g = torch.Generator(device="cuda"); g.manual_seed(1)
k = torch.randint(0, u, (n,), device="cuda", generator=g, dtype=torch.int64) * 7919
rc.MultiKeyGroupBy32([k], 0, None, 2)
torch.cuda.synchronize()
print("---- second call", file=sys.stderr, flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ik, ifk, U = rc.MultiKeyGroupBy32([k], 0, None, 2); e1.record()
torch.cuda.synchronize()
print("total ms", e0.elapsed_time(e1), "uniques", U, file=sys.stderr)
