#!/usr/bin/env python
"""C1 shape (10 M rows, int32 key with 1 K uniques, f64 values): per-call wall time and the hash rounds."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RTB_GB_TRACE"] = "1"
import riptable_b200.rc as rc  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
g = torch.Generator(device="cuda"); g.manual_seed(12345)
k = torch.randint(0, 1000, (n,), device="cuda", generator=g, dtype=torch.int32)
v = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
for _ in range(3):
    ik, ifk, u = rc.MultiKeyGroupBy32([k], 0, None, 2)
torch.cuda.synchronize()
print("---- timed call", file=sys.stderr, flush=True)
t = time.perf_counter(); ik, ifk, u = rc.MultiKeyGroupBy32([k], 0, None, 2); torch.cuda.synchronize(); th = time.perf_counter() - t
ik16 = ik.to(torch.int16)
for _ in range(3):
    rc.GroupByAll32([v], ik16, u, [0, ], [1], [u + 1], 0)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
    s = rc.GroupByAll32([v], ik16, u, [0], [1], [u + 1], 0)
torch.cuda.synchronize(); ts = (time.perf_counter() - t) / 10
t = time.perf_counter()
for _ in range(10):
    s = rc.GroupByAll32([v], ik16, u, [1], [1], [u + 1], 0)
torch.cuda.synchronize(); tm = (time.perf_counter() - t) / 10
print("hash %.0f us, sum %.0f us, mean %.0f us  (wall, device mode)" % (th * 1e6, ts * 1e6, tm * 1e6), file=sys.stderr)
