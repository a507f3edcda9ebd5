#!/usr/bin/env python
"""Host mode with PAGEABLE NumPy inputs (what riptable passes): staged copy through the page-locked ring vs the
driver's own pageable path (RTB_NO_STAGING=1)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
rng = np.random.default_rng(3)
key = rng.integers(0, 100_000, n).astype(np.int64) * 7919
val = rng.random(n)
for mode in ("staged", "driver"):
    os.environ["RTB_NO_STAGING"] = "0" if mode == "staged" else "1"
    for _ in range(2):
        d = rc._to_dev(key)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        d = rc._to_dev(key)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    ok = bool((d.torch().cpu().numpy() == key).all())
    def step_fn():
        ik, ifk, u = rc.MultiKeyGroupBy32([key], 0, None, 2)
        return rc.GroupByAll32([val], ik, u, [0], [1], [u + 1], 0)[0]

    step_fn(); step_fn()  # page-locked result buffers come from torch's caching host allocator after this
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3):
        step_fn()
    torch.cuda.synchronize()
    step = (time.perf_counter() - t) / 3
    print(f"{mode:7s} H2D of {key.nbytes / 1e6:.0f} MB pageable: {dt * 1e3:7.1f} ms = {key.nbytes / dt / 1e9:5.1f} GB/s  correct={ok}   "
          f"groupby+sum host mode: {step * 1e3:7.1f} ms = {n / step / 1e6:6.0f} M rows/s", flush=True)
