"""Generates tests/golden/hotpath_v1.npz: small seeded inputs and the outputs the pinned oracle produces for every
entry point on the path.  The oracle itself is pinned to the reference's docstring vectors and KATs in
tests/test_oracle_golden.py; riptide_cpp cannot be imported here, so these fixtures are oracle outputs,
not riptide_cpp outputs.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import riptide_oracle as orc  # noqa: E402


def build():
    rng = np.random.default_rng(20240601)
    n = 8_000
    g = {}
    k64 = rng.integers(-(2**62), 2**62, 700, dtype=np.int64)[rng.integers(0, 700, n)]
    k32 = rng.integers(0, 90, n).astype(np.int32)
    ks = rng.choice(np.array([b"", b"a", b"ab", b"abcd", b"zz"], dtype="S4"), n)
    kf = np.round(rng.random(n) * 20, 0)
    kf[rng.random(n) < 0.05] = np.nan
    filt = rng.random(n) < 0.7
    v64 = rng.random(n) * 200 - 100
    v64[rng.random(n) < 0.08] = np.nan
    v32 = (rng.random(n) * 50).astype(np.float32)
    vi = rng.integers(-500, 500, n).astype(np.int16)
    vi[rng.random(n) < 0.05] = np.iinfo(np.int16).min
    g.update(k64=k64, k32=k32, ks=ks, kf=kf, filt=filt, v64=v64, v32=v32, vi=vi)

    for name, keys, f in (("a", [k64], None), ("b", [k32, ks], None), ("c", [kf], filt), ("d", [k64, ks, k32], filt)):
        ik, ifk, u = orc.MultiKeyGroupBy32(keys, 0, f, 2)
        g[f"gb_{name}_ikey"], g[f"gb_{name}_ifk"], g[f"gb_{name}_u"] = ik, ifk, np.int64(u)
    ik, (ifk, ucut), u = orc.MultiKeyGroupBy32([k32], 0, None, 2, cutoffs=np.array([2000, 2000, 5_345, n], dtype=np.int64))
    g["gbcut_ikey"], g["gbcut_ifk"], g["gbcut_ucut"], g["gbcut_u"] = ik, ifk, ucut, np.int64(u)

    ik, u = g["gb_b_ikey"], int(g["gb_b_u"])
    for fn in (0, 1, 2, 3, 4, 5, 50, 51, 52, 53, 54, 55):
        for vn, v in (("v64", v64), ("v32", v32), ("vi", vi)):
            g[f"red_{fn}_{vn}"] = orc.GroupByAll32([v], ik, u, [fn], [0], [u + 1], 0)[0]
    g["count"] = orc.BinCount(ik, u + 1)
    g["combine_filter"] = orc.CombineFilter(ik, filt)
    cnt, ig, ifg = orc.BinCount(ik, u + 1, pack=True)
    g["pack_igroup"], g["pack_ifirstgroup"], g["pack_ncount"] = ig, ifg, cnt
    for fn, p in ((100, 0), (102, 0), (101, 1), (101, -2)):
        g[f"pick_{fn}_{p}_v64"] = orc.GroupByAllPack32([v64], ik, ig, ifg, cnt, u, [fn], [1], [u + 1], p)[0]
        g[f"pick_{fn}_{p}_ks"] = orc.GroupByAllPack32([ks], ik, ig, ifg, cnt, u, [fn], [1], [u + 1], p)[0]
    vclean = np.nan_to_num(v64)
    g["cumsum_v64"] = orc.EmaAll32([vclean], ik, u, 300, (0.0, None, None, None))[0]
    g["cumsum_v64_filt"] = orc.EmaAll32([vclean], ik, u, 300, (0.0, None, filt, None))[0]
    g["cumsum_vi"] = orc.EmaAll32([vi], ik, u, 300, (0.0, None, filt, None))[0]
    for mode in (0, 1):
        g[f"ifirst_{mode}"] = orc.MakeiFirst(ik, u, filt, mode)
        g[f"inext_{mode}"] = orc.MakeiNext(ik, u, mode)
    g["lex_k64_kf"] = orc.LexSort32([k64, kf])
    g["lex_ks_k32_desc"] = orc.LexSort32([ks, k32], ascending=[False, True])
    idx = np.flatnonzero(filt).astype(np.int32)
    g["lex_idx"] = idx
    g["lex_kf_index"] = orc.LexSort32([kf], index=idx)
    lik, lifk, lcnt = orc.GroupFromLexSort(g["lex_k64_kf"], [kf, k64], base_index=1)
    g["lexgrp_ikey"], g["lexgrp_ifk"], g["lexgrp_cnt"] = lik, lifk, lcnt
    g["revshuffle"] = orc.ReverseShuffle(g["lex_k64_kf"])
    return g


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_v1.npz")
    np.savez_compressed(out, **build())
    print(out, os.path.getsize(out), "bytes")
