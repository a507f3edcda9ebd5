#!/usr/bin/env python
"""String keys: one S8 / S16 column with few uniques (the shape of a symbol column), per-round trace."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RTB_GB_TRACE"] = "1"
import riptable_b200.rc as rc  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
width = int(sys.argv[2]) if len(sys.argv) > 2 else 16
nu = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
pool = np.random.default_rng(1).integers(65, 91, nu * width, dtype=np.uint8).reshape(nu, width)
g = torch.Generator(device="cuda"); g.manual_seed(1)
idx = torch.randint(0, nu, (n,), device="cuda", generator=g, dtype=torch.int64)
col = rc.DevArray(torch.from_numpy(pool).cuda()[idx].reshape(-1).contiguous(), np.dtype(f"S{width}"), n)
del idx
rc.MultiKeyGroupBy32([col], 0, None, 2); torch.cuda.synchronize()
print("---- second call", file=sys.stderr, flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ik, ifk, u = rc.MultiKeyGroupBy32([col], 0, None, 2); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("S%d, %d rows, %d uniques: %.2f ms = %.1f G rows/s = %.0f GB/s of key bytes" % (width, n, u, ms, n / ms / 1e6, n * width / ms / 1e6), file=sys.stderr)
