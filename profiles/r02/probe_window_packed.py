"""(int64, int32) pairs, 500 M rows: the table-window rule on packed-row tables (RTB_GB_WINDOW_PACKED) by pair count."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc
n = 500_000_000
g = torch.Generator(device="cuda"); g.manual_seed(5)
for pairs in (1_500_000, 2_000_000, 3_000_000):
    a = torch.randint(0, pairs // 50, (n,), device="cuda", generator=g, dtype=torch.int64) * 1_000_003
    b = torch.randint(0, 50, (n,), device="cuda", generator=g, dtype=torch.int32)
    out = {}
    for mode in ("0", "1"):
        os.environ["RTB_GB_WINDOW_PACKED"] = mode
        rc.MultiKeyGroupBy32([a, b], 0, None, 2); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ik, ifk, u = rc.MultiKeyGroupBy32([a, b], 0, None, 2); e1.record(); torch.cuda.synchronize()
        out["window" if mode == "1" else "default"] = round(e0.elapsed_time(e1), 2)
    print(json.dumps({"rows": n, "pairs": int(u), "ms": out}), flush=True)
    del a, b, ik
