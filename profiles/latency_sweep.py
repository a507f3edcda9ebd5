#!/usr/bin/env python
"""Call latency of the two headline entry points across input sizes (the reference's bench_grouping.py grid is
1e5..1e7 rows): device mode (CUDA tensors in/out) and host mode (NumPy in/out), wall clock around synchronised calls."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402

rng = np.random.default_rng(12345)
print(f"{'rows':>10} {'uniques':>8} | {'hash dev us':>11} {'sum dev us':>10} | {'hash host us':>12} {'sum host us':>11}")
for n in (1_000, 100_000, 1_000_000, 10_000_000, 100_000_000):
    for u in (100, 100_000):
        if u > n:
            continue
        key = rng.integers(0, u, n).astype(np.int64) * 7919
        val = rng.random(n)
        kd, vd = torch.from_numpy(key).cuda(), torch.from_numpy(val).cuda()

        def timeit(fn, reps):
            fn(); fn()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(reps):
                out = fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t) / reps * 1e6, out

        reps = 20 if n <= 1_000_000 else 5
        th, (ik, ifk, U) = timeit(lambda: rc.MultiKeyGroupBy32([kd], 0, None, 2), reps)
        ts, _ = timeit(lambda: rc.GroupByAll32([vd], ik, U, [0], [1], [U + 1], 0), reps)
        thh, (ikh, ifkh, Uh) = timeit(lambda: rc.MultiKeyGroupBy32([key], 0, None, 2), reps)
        tsh, _ = timeit(lambda: rc.GroupByAll32([val], ikh, Uh, [0], [1], [Uh + 1], 0), reps)
        print(f"{n:>10} {u:>8} | {th:>11.0f} {ts:>10.0f} | {thh:>12.0f} {tsh:>11.0f}", flush=True)
