"""Per-round trace (RTB_GB_TRACE=1) of the fused group-and-sum call on the C2 recipe at a given row count."""
import os, sys, time
os.environ["RTB_GB_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000_000
u = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
g = torch.Generator(device="cuda"); g.manual_seed(7)
pool = torch.randint(-2**62, 2**62, (u,), device="cuda", generator=g, dtype=torch.int64)
key = pool[torch.randint(0, u, (n,), device="cuda", generator=g)]
val = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
for it in range(3):
    torch.cuda.synchronize()
    print(f"---- call {it}", file=sys.stderr, flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    ik, ifk, U, sums = rc.MultiKeyGroupBySum32([key], val)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("total ms (events)", round(e0.elapsed_time(e1), 3), "wall", round((t1 - t0) * 1e3, 3), "uniques", U, file=sys.stderr)
