"""Per-call latency of rc.MultiKeyGroupBy32 on small inputs (device mode), the sizes riptable's own benchmark grid starts
at (riptable/benchmarks/bench_grouping.py:35-52).  RTB_GB_SMALL=0 gives the round schedule, 1 the single-launch path."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc

out = {}
for n in (10_000, 100_000, 1_000_000):
    for u in (100, 10_000):
        g = torch.Generator(device="cuda"); g.manual_seed(1)
        k = torch.randint(0, u, (n,), device="cuda", generator=g, dtype=torch.int32)
        for _ in range(5):
            rc.MultiKeyGroupBy32([k], 0, None, 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 200
        for _ in range(reps):
            rc.MultiKeyGroupBy32([k], 0, None, 2)
        torch.cuda.synchronize()
        out[f"n{n}_u{u}"] = round((time.perf_counter() - t0) / reps * 1e6, 1)
print(json.dumps({"small_path": os.environ.get("RTB_GB_SMALL", "1"), "us_per_call": out}))
