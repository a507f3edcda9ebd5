"""Replays tests/golden/hotpath_v1.npz through any `rc`-shaped module (the oracle on CPU, riptable_b200.rc on the GPU)."""
import os

import numpy as np

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_v1.npz"))


def _eq(got, exp, what):
    got, exp = np.asarray(got), np.asarray(exp)
    assert got.shape == exp.shape, (what, got.shape, exp.shape)
    if exp.dtype.kind == "f":
        scale = float(np.nanmax(np.abs(exp))) if exp.size and np.isfinite(exp).any() else 1.0
        rtol = 1e-5 if exp.dtype == np.float32 else 1e-12
        np.testing.assert_allclose(got, exp, rtol=rtol, atol=rtol * max(scale, 1.0) * 64, equal_nan=True, err_msg=what)
    else:
        np.testing.assert_array_equal(got, exp, err_msg=what)


def run(rc):
    k64, k32, ks, kf, filt = G["k64"], G["k32"], G["ks"], G["kf"], G["filt"]
    v64, v32, vi = G["v64"], G["v32"], G["vi"]
    n = len(k64)
    for name, keys, f in (("a", [k64], None), ("b", [k32, ks], None), ("c", [kf], filt), ("d", [k64, ks, k32], filt)):
        ik, ifk, u = rc.MultiKeyGroupBy32(keys, 0, f, 2)
        assert u == int(G[f"gb_{name}_u"])
        _eq(ik, G[f"gb_{name}_ikey"], f"gb {name} ikey")
        _eq(ifk, G[f"gb_{name}_ifk"], f"gb {name} ifirstkey")
    ik, (ifk, ucut), u = rc.MultiKeyGroupBy32([k32], 0, None, 2, cutoffs=np.array([2000, 2000, 5_345, n], dtype=np.int64))
    _eq(ik, G["gbcut_ikey"], "cutoffs ikey"); _eq(ifk, G["gbcut_ifk"], "cutoffs ifk"); _eq(ucut, G["gbcut_ucut"], "cutoffs ucut")
    assert u == int(G["gbcut_u"])

    ik, u = G["gb_b_ikey"], int(G["gb_b_u"])
    for fn in (0, 1, 2, 3, 4, 5, 50, 51, 52, 53, 54, 55):
        for vn, v in (("v64", v64), ("v32", v32), ("vi", vi)):
            got = rc.GroupByAll32([v], ik, u, [fn], [0], [u + 1], 0)[0]
            exp = G[f"red_{fn}_{vn}"]
            assert got.dtype == exp.dtype, (fn, vn, got.dtype, exp.dtype)
            _eq(got, exp, f"reduce {fn} {vn}")
    _eq(rc.BinCount(ik, u + 1), G["count"], "bincount")
    _eq(rc.CombineFilter(ik, filt), G["combine_filter"], "combine_filter")
    cnt, ig, ifg = rc.BinCount(ik, u + 1, pack=True)
    _eq(ig, G["pack_igroup"], "igroup"); _eq(ifg, G["pack_ifirstgroup"], "ifirstgroup"); _eq(cnt, G["pack_ncount"], "ncountgroup")
    for fn, p in ((100, 0), (102, 0), (101, 1), (101, -2)):
        got = rc.GroupByAllPack32([v64], ik, ig, ifg, cnt, u, [fn], [1], [u + 1], p)[0]
        _eq(np.asarray(got)[1:].view(np.uint64), G[f"pick_{fn}_{p}_v64"][1:].view(np.uint64), f"pick {fn} {p} f64")
        got = rc.GroupByAllPack32([ks], ik, ig, ifg, cnt, u, [fn], [1], [u + 1], p)[0]
        _eq(np.asarray(got)[1:], G[f"pick_{fn}_{p}_ks"][1:], f"pick {fn} {p} S4")
    vclean = np.nan_to_num(v64)
    _eq(rc.EmaAll32([vclean], ik, u, 300, (0.0, None, None, None))[0], G["cumsum_v64"], "cumsum f64")
    _eq(rc.EmaAll32([vclean], ik, u, 300, (0.0, None, filt, None))[0], G["cumsum_v64_filt"], "cumsum f64 filter")
    _eq(rc.EmaAll32([vi], ik, u, 300, (0.0, None, filt, None))[0], G["cumsum_vi"], "cumsum i16")
    for mode in (0, 1):
        _eq(rc.MakeiFirst(ik, u, filt, mode), G[f"ifirst_{mode}"], f"ifirst {mode}")
        _eq(rc.MakeiNext(ik, u, mode), G[f"inext_{mode}"], f"inext {mode}")
    _eq(rc.LexSort32([k64, kf]), G["lex_k64_kf"], "lexsort")
    _eq(rc.LexSort32([ks, k32], ascending=[False, True]), G["lex_ks_k32_desc"], "lexsort desc")
    _eq(rc.LexSort32([kf], index=G["lex_idx"]), G["lex_kf_index"], "lexsort index")
    lik, lifk, lcnt = rc.GroupFromLexSort(G["lex_k64_kf"], [kf, k64], base_index=1)
    _eq(lik, G["lexgrp_ikey"], "lexgroup ikey"); _eq(lifk, G["lexgrp_ifk"], "lexgroup ifk"); _eq(lcnt, G["lexgrp_cnt"], "lexgroup cnt")
    _eq(rc.ReverseShuffle(G["lex_k64_kf"]), G["revshuffle"], "reverse shuffle")


def run_v2(rc):
    """The SURVEY 8f rows against tests/golden/hotpath_v2.npz (inputs shared with v1)."""
    H = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_v2.npz"))
    k64, k32, ks, filt = G["k64"], G["k32"], G["ks"], G["filt"]
    v64, v32, vi = G["v64"], G["v32"], G["vi"]
    n = len(k64)
    ik, u = G["gb_b_ikey"], int(G["gb_b_u"])
    cnt, ig, ifg = G["pack_ncount"], G["pack_igroup"], G["pack_ifirstgroup"]
    reset, t = H["reset"], H["time"]

    def same(got, exp, what):
        assert np.asarray(got).dtype == exp.dtype, (what, np.asarray(got).dtype, exp.dtype)
        _eq(got, exp, what)

    f, i = rc.IsMember32(k64, H["probe64"])
    same(f, H["ism_found"], "ismember found"); same(i, H["ism_idx"], "ismember index")
    f, i = rc.IsMemberCategorical(k32, np.arange(50, 130, dtype=np.int32))
    same(f, H["ismcat_found"], "ismember categorical found"); same(i, H["ismcat_idx"], "ismember categorical index")
    f, i = rc.MultiKeyIsMember32(([k32, ks],), ([k32[:500], ks[:500]],))
    same(f, H["mkism_found"], "multikey ismember found"); same(i, H["mkism_idx"], "multikey ismember index")
    for fn in (302, 306, 307, 308, 309):
        for vn, v in (("v64", v64), ("v32", v32), ("vi", vi)):
            vv = np.clip(np.nan_to_num(v, nan=1.0) / 50 + 1, 0.5, 1.5).astype(v.dtype) if fn == 302 and v.dtype.kind == "f" else v
            same(rc.EmaAll32([vv], ik, u, fn, (0.0, None, filt, reset))[0], H[f"cum_{fn}_{vn}"], f"cum op {fn} {vn}")
    for fn, rate in ((301, 0.01), (304, 0.05), (305, 0.8)):
        same(rc.EmaAll32([np.nan_to_num(v64)], ik, u, fn, (rate, t, filt, reset))[0], H[f"ema_{fn}_v64"], f"ema {fn} f64")
        same(rc.EmaAll32([vi], ik, u, fn, (rate, t, None, None))[0], H[f"ema_{fn}_vi"], f"ema {fn} i16")
    for fn, p in ((200, 3), (201, 5), (202, 1), (202, -2), (203, 2), (203, -1), (204, 1), (204, -1), (205, 4), (206, 4)):
        for vn, v in (("v64", v64), ("vi", vi)):
            same(rc.GroupByAllPack32([v], ik, ig, ifg, cnt, u, [fn], [0], [u + 1], p)[0], H[f"roll_{fn}_{p}_{vn}"], f"rolling {fn} {p} {vn}")
    same(rc.GroupByAllPack32([ks], ik, ig, ifg, cnt, u, [203], [0], [u + 1], 1)[0], H["roll_203_1_ks"], "shift S4")
    for key in H.files:
        if key.startswith("quant_"):
            _, fp, vn = key.split("_")
            v = {"v64": v64, "v32": v32, "vi": vi}[vn]
            same(rc.GroupByAllPack32([v], ik, ig, ifg, cnt, u, [106], [1], [u + 1], int(fp))[0], H[key], key)
    same(rc.GroupByAllPack32([v64], ik, ig, ifg, cnt, u, [103], [1], [u + 1], 0)[0], H["median_v64"], "median")
    same(rc.GroupByAllPack32([H["vm"]], ik, ig, ifg, cnt, u, [104], [1], [u + 1], 0)[0], H["mode_vm"], "mode i32")
    same(rc.GroupByAllPack32([vi], ik, ig, ifg, cnt, u, [104], [1], [u + 1], 0)[0], H["mode_vi"], "mode i16")
    same(rc.ReIndexGroups(H["reidx_ikey"], H["reidx_uik"], H["reidx_ucut"], H["reidx_cuts"]), H["reidx_out"], "reindex groups")
    k1 = (k32 % 7 + 1).astype(np.int8)
    k2 = (H["vm"] + 1).astype(np.int8)
    a_ik, a_ifk, a_u = rc.CombineAccum1Filter(k1, 7, filt)
    same(a_ik, H["acc1_ikey"], "accum1 ikey"); same(a_ifk, H["acc1_ifk"], "accum1 ifirstkey"); assert int(a_u) == int(H["acc1_u"])
    b_ik, b_cnt = rc.CombineAccum2Filter(k1, k2, 7, 6, filt)
    same(b_ik, H["acc2_ikey"], "accum2 ikey"); same(b_cnt, H["acc2_cnt"], "accum2 counts")
