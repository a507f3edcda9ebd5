"""Per-round trace (RTB_GB_TRACE=1) of the C5 grouping: 1 B rows of an int64 key with 100 M uniques."""
import os, sys
os.environ["RTB_GB_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
u = n // 10
g = torch.Generator(device="cuda"); g.manual_seed(11)
key = torch.randint(0, u, (n,), device="cuda", generator=g, dtype=torch.int64) * 7_919
for it in range(2):
    torch.cuda.synchronize()
    print(f"---- call {it}", file=sys.stderr, flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ik, ifk, U = rc.MultiKeyGroupBy32([key], 0, None, 2)
    e1.record(); torch.cuda.synchronize()
    print("total ms", round(e0.elapsed_time(e1), 3), "uniques", U, file=sys.stderr)
    del ik, ifk
