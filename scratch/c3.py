import sys, torch, numpy as np
sys.path.insert(0, '.')
import riptable_b200.rc as rc
n = int(sys.argv[1]); nk1 = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
g = torch.Generator(device='cuda'); g.manual_seed(3)
k1 = torch.randint(0, nk1, (n,), device='cuda', generator=g, dtype=torch.int64) * 1_000_003
pool = np.frombuffer(np.random.default_rng(1).integers(65, 91, 1000 * 16, dtype=np.uint8).tobytes(), dtype="S16")
sidx = torch.randint(0, 1000, (n,), device='cuda', generator=g, dtype=torch.int64)
s16 = rc.DevArray(torch.from_numpy(pool.view(np.uint8).reshape(1000, 16).copy()).cuda()[sidx].reshape(-1).contiguous(), np.dtype("S16"), n)
del sidx
k3 = torch.randint(0, 5, (n,), device='cuda', generator=g, dtype=torch.int32)
import os
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ik, ifk, u = rc.MultiKeyGroupBy32([k1, s16, k3], 0, None, 2); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"C3 multikey (int64,S16,int32) rows {n} uniques {u}: {ms:.1f} ms, {n/ms/1e6:.2f} G rows/s, alg {n*32/ms/1e6:.0f} GB/s", flush=True)
