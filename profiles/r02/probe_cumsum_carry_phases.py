"""One payload-carrying GB_CUMSUM call (268 M rows, 1 M bins) for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 268_435_456
g = torch.Generator(device="cuda"); g.manual_seed(3)
val = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
ik = torch.randint(1, 1_000_001, (n,), device="cuda", generator=g, dtype=torch.int32)
for mode in ("1", "0"):
    os.environ["RTB_CUMSUM_CARRY"] = mode
    r = rc.EmaAll32([val], ik, 1_000_000, 300, (0.0, None, None, None))[0]
    torch.cuda.synchronize()
