#!/usr/bin/env python
"""rc.IsMember32: 1 B int64 rows looked up in a 1 M-row set (half of the probes hit)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402
g = torch.Generator(device="cuda"); g.manual_seed(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
uv = torch.randint(-(2**62), 2**62, (1_000_000,), dtype=torch.int64, device="cuda", generator=g)
k = torch.empty(n, dtype=torch.int64, device="cuda")
for s in range(0, n, 1 << 27):
    e = min(n, s + (1 << 27))
    torch.index_select(uv, 0, torch.randint(0, 1_000_000, (e - s,), device="cuda", generator=g), out=k[s:e])
probe = torch.cat([uv[:500_000], torch.randint(-(2**62), 2**62, (500_000,), dtype=torch.int64, device="cuda", generator=g)])
rc.IsMember32(k, probe); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    m, idx = rc.IsMember32(k, probe)
e1.record(); torch.cuda.synchronize()
print("IsMember32 %d rows in 1M set: %.2f ms per call, found frac %.3f" % (n, e0.elapsed_time(e1) / 3, m.float().mean().item()))
