"""One blocked GB_CUMSUM call per bin count, for an ncu launch list (kernel time per phase)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 268_435_456
g = torch.Generator(device="cuda"); g.manual_seed(3)
val = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
for u in (100, 3_000, 10_000):
    ik = torch.randint(1, u + 1, (n,), device="cuda", generator=g, dtype=torch.int32)
    r = rc.EmaAll32([val], ik, u, 300, (0.0, None, None, None))[0]
    torch.cuda.synchronize()
    del ik, r
