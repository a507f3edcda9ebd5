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


def build_v2():
    """hotpath_v2.npz: the SURVEY 8f rows (ismember, cumulative / rolling / quantile / mode / ema functions, ReIndexGroups,
    CombineAccum1/2Filter) on the inputs of v1."""
    v1 = build()
    rng = np.random.default_rng(20240602)
    g = {}
    k64, k32, ks, kf, filt, v64, v32, vi = (v1[x] for x in ("k64", "k32", "ks", "kf", "filt", "v64", "v32", "vi"))
    n = len(k64)
    ik, u = v1["gb_b_ikey"], int(v1["gb_b_u"])
    cnt, ig, ifg = v1["pack_ncount"], v1["pack_igroup"], v1["pack_ifirstgroup"]
    reset = rng.random(n) < 0.02
    t = np.cumsum(rng.integers(0, 4, n)).astype(np.int64)
    g.update(reset=reset, time=t)
    probe64 = np.concatenate([k64[:300], rng.integers(-(2**62), 2**62, 200, dtype=np.int64)])
    g["probe64"] = probe64
    g["ism_found"], g["ism_idx"] = orc.IsMember32(k64, probe64)
    g["ismcat_found"], g["ismcat_idx"] = orc.IsMemberCategorical(k32, np.arange(50, 130, dtype=np.int32))
    g["mkism_found"], g["mkism_idx"] = orc.MultiKeyIsMember32(([k32, ks],), ([k32[:500], ks[:500]],))
    for fn in (302, 306, 307, 308, 309):
        for vn, v in (("v64", v64), ("v32", v32), ("vi", vi)):
            vv = np.clip(np.nan_to_num(v, nan=1.0) / 50 + 1, 0.5, 1.5).astype(v.dtype) if fn == 302 and v.dtype.kind == "f" else v
            g[f"cum_{fn}_{vn}"] = orc.EmaAll32([vv], ik, u, fn, (0.0, None, filt, reset))[0]
    for fn, rate in ((301, 0.01), (304, 0.05), (305, 0.8)):
        g[f"ema_{fn}_v64"] = orc.EmaAll32([np.nan_to_num(v64)], ik, u, fn, (rate, t, filt, reset))[0]
        g[f"ema_{fn}_vi"] = orc.EmaAll32([vi], ik, u, fn, (rate, t, None, None))[0]
    for fn, p in ((200, 3), (201, 5), (202, 1), (202, -2), (203, 2), (203, -1), (204, 1), (204, -1), (205, 4), (206, 4)):
        for vn, v in (("v64", v64), ("vi", vi)):
            g[f"roll_{fn}_{p}_{vn}"] = orc.GroupByAllPack32([v], ik, ig, ifg, cnt, u, [fn], [0], [u + 1], p)[0]
    g["roll_203_1_ks"] = orc.GroupByAllPack32([ks], ik, ig, ifg, cnt, u, [203], [0], [u + 1], 1)[0]
    for q, isnan in ((0.5, True), (0.5, False), (0.1, True), (0.9, False), (0.0, True), (1.0, True)):
        fp = orc.quantile_func_param(q, isnan)
        for vn, v in (("v64", v64), ("v32", v32), ("vi", vi)):
            g[f"quant_{fp}_{vn}"] = orc.GroupByAllPack32([v], ik, ig, ifg, cnt, u, [106], [1], [u + 1], fp)[0]
    vm = rng.integers(0, 6, n).astype(np.int32)
    g["vm"] = vm
    g["median_v64"] = orc.GroupByAllPack32([v64], ik, ig, ifg, cnt, u, [103], [1], [u + 1], 0)[0]
    g["mode_vm"] = orc.GroupByAllPack32([vm], ik, ig, ifg, cnt, u, [104], [1], [u + 1], 0)[0]
    g["mode_vi"] = orc.GroupByAllPack32([vi], ik, ig, ifg, cnt, u, [104], [1], [u + 1], 0)[0]
    cuts = np.array([1500, 1500, 4000, n], dtype=np.int64)
    pik, (pifk, ucut), pu = orc.MultiKeyGroupBy32([k32], 0, None, 2, cutoffs=cuts)
    uik = rng.integers(1, 500, int(ucut[-1])).astype(np.int32)
    g.update(reidx_cuts=cuts, reidx_ucut=ucut, reidx_uik=uik, reidx_ikey=pik)
    g["reidx_out"] = orc.ReIndexGroups(pik, uik, ucut, cuts)
    k1 = (k32 % 7 + 1).astype(np.int8)
    k2 = (vm + 1).astype(np.int8)
    g["acc1_ikey"], g["acc1_ifk"], g["acc1_u"] = orc.CombineAccum1Filter(k1, 7, filt)
    g["acc2_ikey"], g["acc2_cnt"] = orc.CombineAccum2Filter(k1, k2, 7, 6, filt)
    return g


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, fn in (("hotpath_v1.npz", build), ("hotpath_v2.npz", build_v2)):
        out = os.path.join(here, name)
        np.savez_compressed(out, **fn())
        print(out, os.path.getsize(out), "bytes")
