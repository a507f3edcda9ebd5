"""GB_CUMSUM of 1 B float64 rows above the blocked form's bin limit: the payload-carrying sort against the gather / scatter
form (RTB_CUMSUM_CARRY=0), at 1 M and 100 M bins."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import riptable_b200.rc as rc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(3)
val = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


for u in (100_000, 1_000_000, n // 10):
    ik = torch.randint(1, u + 1, (n,), device="cuda", generator=g, dtype=torch.int32)
    out = {}
    for mode in ("1", "0"):
        os.environ["RTB_CUMSUM_CARRY"] = mode
        ms, r = timed(lambda: rc.EmaAll32([val], ik, u, 300, (0.0, None, None, None))[0])
        out["carried" if mode == "1" else "gather_scatter"] = round(ms, 2)
        if mode == "1":
            keep = r
        else:
            err = float((keep - r).abs().max())
    print(json.dumps({"rows": n, "bins": u, "ms": out, "max_abs_diff_between_forms": err}), flush=True)
    del ik, keep, r
    torch.cuda.empty_cache()
