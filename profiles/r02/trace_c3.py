"""Per-round trace (RTB_GB_TRACE=1) of the C3 multi-key grouping: 500 M rows of (int64, S16, int32), about 5e7 tuples."""
import os, sys
os.environ["RTB_GB_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import riptable_b200.rc as rc
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000_000
g = torch.Generator(device="cuda"); g.manual_seed(12345)
k1 = torch.randint(0, 10_000, (n3,), device="cuda", generator=g, dtype=torch.int64) * 1_000_003
pool = np.frombuffer(np.random.default_rng(1).integers(65, 91, 1000 * 16, dtype=np.uint8).tobytes(), dtype="S16")
sidx = torch.randint(0, 1000, (n3,), device="cuda", generator=g, dtype=torch.int64)
s16 = rc.DevArray(torch.from_numpy(pool.view(np.uint8).reshape(1000, 16).copy()).cuda()[sidx].reshape(-1).contiguous(), np.dtype("S16"), n3)
del sidx
k3 = torch.randint(0, 5, (n3,), device="cuda", generator=g, dtype=torch.int32)
rc.MultiKeyGroupBy32([k1, s16, k3], 0, None, 2)
torch.cuda.synchronize()
print("---- second call", file=sys.stderr, flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ik, ifk, U = rc.MultiKeyGroupBy32([k1, s16, k3], 0, None, 2)
e1.record(); torch.cuda.synchronize()
print("total ms", e0.elapsed_time(e1), "uniques", U, file=sys.stderr)
