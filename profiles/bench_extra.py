#!/usr/bin/env python
"""Secondary measurements (not the headline bench): the other BASELINE.json configs on one B200, device mode, CUDA
events, 1 warm-up + 3 timed.  Prints one JSON line per operation with rows/s and algorithmic GB/s."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import riptable_b200.rc as rc  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def line(name, n, ms, alg_bytes_per_row, **kw):
    print(json.dumps({"op": name, "rows": n, "ms": round(ms, 3), "rows_per_s": n / ms * 1e3,
                      "alg_bytes_per_row": alg_bytes_per_row, "alg_GBps": n * alg_bytes_per_row / ms / 1e6, **kw}), flush=True)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(12345)

    def keys(u, dtype=torch.int64):
        uv = torch.randint(-(2**62), 2**62, (u,), dtype=torch.int64, device="cuda", generator=g)
        k = torch.empty(n, dtype=torch.int64, device="cuda")
        for s in range(0, n, 1 << 27):
            e = min(n, s + (1 << 27))
            torch.index_select(uv, 0, torch.randint(0, u, (e - s,), device="cuda", generator=g), out=k[s:e])
        return k

    # C1-like: int32 key with 1K uniques, f64 sum + mean on int16 iKey
    k32 = torch.randint(0, 1000, (n,), device="cuda", generator=g, dtype=torch.int32)
    ms, (ik, ifk, u) = timed(lambda: rc.MultiKeyGroupBy32([k32], 0, None, 2))
    line("C1 hash int32 key 1K uniques", n, ms, 8, uniques=u)
    ik16 = ik.to(torch.int16)
    val = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
    ms, _ = timed(lambda: rc.GroupByAll32([val], ik16, u, [0], [1], [u + 1], 0))
    line("C1 sum f64 by int16 iKey (smem-privatised)", n, ms, 10)
    ms, _ = timed(lambda: rc.GroupByAll32([val], ik16, u, [1], [1], [u + 1], 0))
    line("C1 mean f64 by int16 iKey", n, ms, 10)
    ms, _ = timed(lambda: rc.EmaAll32([val], ik16, u, 300, (0.0, None, None, None)), reps=2)
    line("C1 cumsum f64 by int16 iKey (blocked form, no sort)", n, ms, 18)
    del k32, ik16

    # C2: int64 key with 1M uniques; the other reductions the config names (sum f64 is the headline bench)
    k = keys(1_000_000)
    ms, (ik, ifk, u) = timed(lambda: rc.MultiKeyGroupBy32([k], 0, None, 2))
    line("C2 hash int64 key 1M uniques", n, ms, 12, uniques=u)
    del k
    val32 = val.to(torch.float32)
    for name, v, fn, b in (("sum f64", val, 0, 12), ("sum f32", val32, 0, 8), ("min f64", val, 2, 12), ("max f32", val32, 3, 8),
                           ("mean f64", val, 1, 12), ("nansum f64", val, 50, 12), ("var f64 (two passes)", val, 4, 12)):
        ms, _ = timed(lambda: rc.GroupByAll32([v], ik, u, [fn], [1], [u + 1], 0))
        line("C2 " + name, n, ms, b)
    ms, _ = timed(lambda: rc.BinCount(ik, u + 1))
    line("C2 count (BinCount)", n, ms, 4)
    del val32

    # SURVEY 8f rows on the C2 grouping (1M bins): algorithmic bytes = inputs read once + outputs written once
    ms, (cnt, ig, ifg) = timed(lambda: rc.BinCount(ik, u + 1, pack=True), reps=2)
    line("8f C2 pack (BinCount pack=True, 20-bit bins)", n, ms, 8)
    t64 = torch.arange(n, device="cuda", dtype=torch.int64)
    for name, fn, b in (("cumsum f64", lambda: rc.EmaAll32([val], ik, u, 300, (0.0, None, None, None)), 20),
                        ("cummax f64 (skipna)", lambda: rc.EmaAll32([val], ik, u, 306, (0.0, None, None, None)), 20),
                        ("ema_normal f64, int64 time", lambda: rc.EmaAll32([val], ik, u, 304, (1e-6, t64, None, None)), 28),
                        ("rolling_sum f64 window 3 (packed layout given)", lambda: rc.GroupByAllPack32([val], ik, ig, ifg, cnt, u, [200], [0], [u + 1], 3), 24),
                        ("shift f64 by 1 (packed layout given)", lambda: rc.GroupByAllPack32([val], ik, ig, ifg, cnt, u, [203], [0], [u + 1], 1), 24),
                        ("median f64 (packed layout given)", lambda: rc.GroupByAllPack32([val], ik, ig, ifg, cnt, u, [103], [1], [u + 1], 0), 12)):
        ms, _ = timed(fn, reps=2)
        line("8f C2 " + name, n, ms, b)
    del t64, ig, ifg, cnt
    probe = torch.randint(-(2**62), 2**62, (1_000_000,), dtype=torch.int64, device="cuda", generator=g)
    kk = keys(1_000_000)
    ms, _ = timed(lambda: rc.IsMember32(kk, probe), reps=2)
    line("8f IsMember32: 1B int64 rows in a 1M-row set", n, ms, 8 + 1 + 4)
    del kk, probe, ik

    # C5: 100M uniques (scaled with n), cumsum / pack / first / last
    u5 = max(n // 10, 1)
    k = keys(u5)
    ms, (ik, ifk, u) = timed(lambda: rc.MultiKeyGroupBy32([k], 0, None, 2))
    line("C5 hash int64 key, N/10 uniques", n, ms, 12, uniques=u)
    del k
    ms, (cnt, ig, ifg) = timed(lambda: rc.BinCount(ik, u + 1, pack=True))
    line("C5 pack (BinCount pack=True)", n, ms, 8)
    ms, _ = timed(lambda: rc.GroupByAllPack32([val], ik, ig, ifg, cnt, u, [100], [1], [u + 1], 0))
    line("C5 first", n, ms, (4 + 8) * u / n)
    ms, _ = timed(lambda: rc.EmaAll32([val], ik, u, 300, (0.0, None, None, None)))
    line("C5 cumsum f64", n, ms, 20)
    ms, _ = timed(lambda: rc.GroupByAll32([val], ik, u, [0], [1], [u + 1], 0))
    line("C5 sum f64 (global atomics, N/10 bins)", n, ms, 12)
    del ig, ifg, cnt, ik

    # C3: multikey (int64, S16, int32), N/10 distinct tuples, at half the rows (BASELINE: 500 M rows, 50 M uniques)
    n3 = n // 2
    k1 = torch.randint(0, 10_000, (n3,), device="cuda", generator=g, dtype=torch.int64) * 1_000_003
    pool = np.frombuffer(np.random.default_rng(1).integers(65, 91, 1000 * 16, dtype=np.uint8).tobytes(), dtype="S16")
    sidx = torch.randint(0, 1000, (n3,), device="cuda", generator=g, dtype=torch.int64)
    s16 = rc.DevArray(torch.from_numpy(pool.view(np.uint8).reshape(1000, 16).copy()).cuda()[sidx].reshape(-1).contiguous(),
                      np.dtype("S16"), n3)
    del sidx
    k3 = torch.randint(0, 5, (n3,), device="cuda", generator=g, dtype=torch.int32)
    rc.MultiKeyGroupBy32([k1, s16, k3], 0, None, 2)  # lets the allocator settle on the large workspace
    ms, (ik3, ifk3, u3) = timed(lambda: rc.MultiKeyGroupBy32([k1, s16, k3], 0, None, 2), reps=2)
    print(json.dumps({"op": "C3 hash multikey (int64, S16, int32)", "rows": n3, "ms": round(ms, 3), "rows_per_s": n3 / ms * 1e3,
                      "alg_bytes_per_row": 32, "alg_GBps": n3 * 32 / ms / 1e6, "uniques": u3}), flush=True)
    v3 = val[:n3]
    ms, _ = timed(lambda: rc.GroupByAll32([v3], ik3, u3, [51], [1], [u3 + 1], 0))
    print(json.dumps({"op": "C3 nanmean f64", "rows": n3, "ms": round(ms, 3), "rows_per_s": n3 / ms * 1e3}), flush=True)
    ms, _ = timed(lambda: rc.GroupByAll32([v3], ik3, u3, [55], [1], [u3 + 1], 0))
    print(json.dumps({"op": "C3 nanstd f64 (two passes)", "rows": n3, "ms": round(ms, 3), "rows_per_s": n3 / ms * 1e3}), flush=True)
    del k1, s16, k3, ik3, ifk3

    # C4: lexsort on (int64 1e3 values, f64 with NaN, int32 1e6 values), int64 primary; then GroupFromLexSort
    a = torch.randint(0, 1000, (n,), device="cuda", generator=g, dtype=torch.int64)
    f = val
    c = torch.randint(0, 1_000_000, (n,), device="cuda", generator=g, dtype=torch.int32)
    ms, perm = timed(lambda: rc.LexSort32([c, f, a]), reps=2)
    line("C4 lexsort 3 keys (int32, f64, int64)", n, ms, 24)
    ms, _ = timed(lambda: rc.GroupFromLexSort(perm, [a, f, c], base_index=1), reps=2)
    line("C4 GroupFromLexSort", n, ms, 28)
    ms, _ = timed(lambda: rc.LexSort32([c]), reps=2)
    line("lexsort 1 key int32 (1e6 values)", n, ms, 8)


if __name__ == "__main__":
    main()
