#!/usr/bin/env python
"""The reference's published lexsort benchmark (docs/source/benchmarking.rst:48-77; riptable/benchmarks/bench_numpy.py:
338-358): lexsort((arr, arr)) on np.arange(n) of int32 / float32, n = 100 .. 10 M; the published riptable figures are
10 M rows: min 39.8 ms, median 92.8 ms on 8 threads; np.lexsort 105.8 ms.  Here: riptable_b200.rc.LexSort32, NumPy in /
NumPy out and CUDA in / CUDA out, 1 warm-up + 5 timed, wall clock."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402


def stats_us(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e6)
    return round(min(ts), 1), round(float(np.median(ts)), 1)


for dt in (np.int32, np.float32):
    for n in (100, 1000, 10_000, 100_000, 1_000_000, 10_000_000, 100_000_000):
        for order in ("ordered", "reversed"):
            arr = np.arange(n, dtype=dt)
            if order == "reversed":
                arr = np.flip(arr).copy()
            hmin, hmed = stats_us(lambda: rc.LexSort32([arr, arr]))
            d = torch.from_numpy(arr).cuda()
            dmin, dmed = stats_us(lambda: rc.LexSort32([d, d]))
            t = time.perf_counter()
            ref = np.lexsort((arr, arr)) if n <= 10_000_000 else None
            tnp = (time.perf_counter() - t) * 1e6 if ref is not None else None
            if ref is not None:
                assert np.array_equal(rc.LexSort32([arr, arr]), ref)
            print(json.dumps({"dtype": np.dtype(dt).name, "rows": n, "order": order, "host_min_us": hmin, "host_median_us": hmed,
                              "device_min_us": dmin, "device_median_us": dmed, "np_lexsort_us_here": None if tnp is None else round(tnp, 1)}), flush=True)
