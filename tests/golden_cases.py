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
