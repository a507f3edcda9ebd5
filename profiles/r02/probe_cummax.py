"""GB_CUMNANMAX of 1 B float64 rows: the blocked (sort-free) form against the packed form, by bin count."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(3)
val = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
for u in (100, 10_000):
    ik = torch.randint(1, u + 1, (n,), device="cuda", generator=g, dtype=torch.int32)
    out = {}
    for mode in ("1", "0"):
        os.environ["RTB_CUMSUM_BLOCKED"] = mode
        f = lambda: rc.EmaAll32([val], ik, u, 306, (0.0, None, None, None))[0]
        r = f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(); r = f(); e1.record(); torch.cuda.synchronize()
        out["blocked" if mode == "1" else "sorted"] = round(e0.elapsed_time(e1) / 2, 2)
        if mode == "1":
            keep = r
        else:
            same = bool((keep == r).all())
    print(json.dumps({"rows": n, "bins": u, "ms": out, "identical": same}), flush=True)
    del ik, keep, r
