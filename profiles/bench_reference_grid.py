#!/usr/bin/env python
"""The reference's own bench_groupbyhash grid (riptable/benchmarks/bench_grouping.py:26-108: dtypes int16, int32, S4, S10,
U8; 100 and 10 000 unique keys; 1e4 .. 1e8 rows; 1 warm-up + 5 timed, wall clock) through riptable_b200.rc:
NumPy in / NumPy out (host mode, what riptable calls) and CUDA in / CUDA out (device mode).  One JSON line per cell."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402


def rand_keyarray(rng, length, dtype, unique_count):
    # riptable/benchmarks/rand_data.py:92-172
    dtype = np.dtype(dtype)
    if dtype.kind in "iu":
        uniques = np.arange(0, unique_count, dtype=dtype)
    elif dtype.kind == "S":
        uniques = rng.integers(65, 90, size=unique_count * dtype.itemsize, dtype=np.int8, endpoint=True).view(dtype)
    else:
        uniques = rng.integers(65, 90, size=unique_count * (dtype.itemsize // 4), dtype=np.int32, endpoint=True).view(dtype)
    return uniques[rng.integers(0, unique_count, length)]


def median_us(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e6


def main():
    max_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    for dtype in (np.int16, np.int32, np.dtype("S4"), np.dtype("S10"), np.dtype("<U8")):
        for uniq in (100, 10_000):
            for rows in (10_000, 100_000, 1_000_000, 10_000_000, 100_000_000):
                if rows > max_rows:
                    continue
                rng = np.random.default_rng(12345)
                key = rand_keyarray(rng, rows, dtype, uniq)
                host = median_us(lambda: rc.MultiKeyGroupBy32([key], 0, None, 2))
                dkey = rc._to_dev(key)
                dev = median_us(lambda: rc.MultiKeyGroupBy32([dkey], 0, None, 2))
                print(json.dumps({"dtype": np.dtype(dtype).str, "uniques": uniq, "rows": rows, "host_us": round(host, 1),
                                  "device_us": round(dev, 1), "device_Grows_s": round(rows / dev / 1e3, 2)}), flush=True)
                del dkey, key


if __name__ == "__main__":
    main()
