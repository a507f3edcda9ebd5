#!/usr/bin/env python
"""One LexSort32 of a single int32 key (1e6 distinct values) -- the radix passes alone, for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
g = torch.Generator(device="cuda"); g.manual_seed(1)
c = torch.randint(0, 1_000_000, (n,), device="cuda", generator=g, dtype=torch.int32)
rc.LexSort32([c]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); p = rc.LexSort32([c]); e1.record(); torch.cuda.synchronize()
print("rows", n, "ms", e0.elapsed_time(e1), "G rows/s", n / e0.elapsed_time(e1) / 1e6)
