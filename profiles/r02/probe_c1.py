"""Low-cardinality grouping (shared-memory table) at 1 B rows: int32 / int64 / int16 keys with 1 K uniques."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
for dt, b in ((torch.int16, 2), (torch.int32, 4), (torch.int64, 8)):
    k = torch.randint(0, 1000, (n,), device="cuda", generator=g, dtype=dt)
    rc.MultiKeyGroupBy32([k], 0, None, 2); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ik, ifk, u = rc.MultiKeyGroupBy32([k], 0, None, 2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(json.dumps({"key": str(dt), "rows": n, "uniques": u, "ms": round(ms, 3), "GBps": round(n * (b + 4) / ms / 1e6, 1)}), flush=True)
    del k, ik
