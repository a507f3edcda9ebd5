"""GPU parity on many small random cases around the kernels' tile boundaries (warp tiles of 128 rows, radix tiles of
4096, scan tiles of 2048): every entry point against the oracle, bit-exact for integer results."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import riptide_oracle as orc

SIZES = [0, 1, 2, 31, 32, 33, 127, 128, 129, 255, 257, 2047, 2048, 2049, 4095, 4096, 4097, 8191, 8193, 12_289]


@pytest.fixture(scope="module")
def rc():
    import riptable_b200.rc as rc

    return rc


def _values(rng, n, dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        v = rng.normal(0, 50, n).astype(dt)
        v[rng.random(n) < 0.1] = np.nan
    elif dt.kind == "b":
        v = rng.integers(0, 2, n).astype(dt)
    else:
        v = rng.integers(-100 if dt.kind == "i" else 0, 100, n).astype(dt)
        v[rng.random(n) < 0.1] = orc.invalid_for(dt)
    return v


def _close(got, exp, what):
    assert got.dtype == exp.dtype, (what, got.dtype, exp.dtype)
    if exp.dtype.kind == "f":
        tol = 2e-5 if exp.dtype == np.float32 else 1e-11
        scale = float(np.nanmax(np.abs(exp[np.isfinite(exp)]))) + 1.0 if np.isfinite(exp).any() else 1.0
        np.testing.assert_allclose(got, exp, rtol=tol, atol=tol * scale, equal_nan=True, err_msg=what)
    else:
        np.testing.assert_array_equal(got, exp, err_msg=what)


@pytest.mark.parametrize("n", SIZES)
def test_fuzz_grouping_and_reductions(rc, n):
    rng = np.random.default_rng(1000 + n)
    for nuniq in (1, 7, max(1, n // 3)):
        k = rng.integers(-3, nuniq, n).astype(rng.choice([np.int8, np.int32, np.int64]) if nuniq < 100 else np.int64)
        filt = rng.random(n) < 0.8 if rng.random() < 0.5 else None
        ik, ifk, u = rc.MultiKeyGroupBy32([k], 0, filt, 2)
        eik, eifk, eu = orc.MultiKeyGroupBy32([k], 0, filt, 2)
        assert u == eu
        np.testing.assert_array_equal(ik, eik)
        np.testing.assert_array_equal(ifk, eifk)
        s = rng.choice(np.array(["", "x", "yy", "zzzz"]), n)
        ik2, ifk2, u2 = rc.MultiKeyGroupBy32([k, s], 0, filt, 2)
        eik2, eifk2, eu2 = orc.MultiKeyGroupBy32([k, s], 0, filt, 2)
        assert u2 == eu2
        np.testing.assert_array_equal(ik2, eik2)
        np.testing.assert_array_equal(ifk2, eifk2)
        if n == 0:
            continue
        np.testing.assert_array_equal(rc.BinCount(eik, eu + 1), orc.BinCount(eik, eu + 1))
        for dt in (np.float64, np.float32, np.int16, np.uint64, np.bool_):
            v = _values(rng, n, dt)
            for fn in (0, 1, 2, 3, 4, 50, 51, 52, 53, 55):
                exp = orc.GroupByAll32([v], eik, eu, [fn], [0], [eu + 1], 0)[0]
                got = rc.GroupByAll32([v], eik, eu, [fn], [0], [eu + 1], 0)[0]
                if exp is None:
                    assert got is None
                    continue
                _close(got, exp, f"n {n} func {fn} dtype {np.dtype(dt)}")


@pytest.mark.parametrize("n", [s for s in SIZES if s > 0])
def test_fuzz_packed_and_row_shaped(rc, n):
    rng = np.random.default_rng(2000 + n)
    nb = int(rng.integers(1, max(2, n // 2 + 1)))
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    filt = rng.random(n) < 0.7
    reset = rng.random(n) < 0.05
    cnt, ig, ifg = rc.BinCount(ikey, nb + 1, pack=True)
    ecnt, eig, eifg = orc.BinCount(ikey, nb + 1, pack=True)
    np.testing.assert_array_equal(cnt, ecnt); np.testing.assert_array_equal(ig, eig); np.testing.assert_array_equal(ifg, eifg)
    t = np.cumsum(rng.integers(0, 3, n)).astype(np.int64)
    for dt in (np.float64, np.int32, np.uint8):
        v = _values(rng, n, dt)
        for fn, p in ((100, 0), (102, 0), (101, 1), (101, -1)):
            _close(rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [fn], [1], [nb + 1], p)[0][1:],
                   orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [fn], [1], [nb + 1], p)[0][1:], f"pick {fn} {p}")
        for fn in (300, 306, 307, 308, 309):
            vv = np.nan_to_num(v) if fn == 300 else v
            _close(rc.EmaAll32([vv], ikey, nb, fn, (0.0, None, filt, reset))[0], orc.EmaAll32([vv], ikey, nb, fn, (0.0, None, filt, reset))[0], f"cum {fn}")
        for fn, rate in ((301, 0.02), (304, 0.1), (305, 0.7)):
            vv = np.nan_to_num(v.astype(np.float64)) if np.dtype(dt).kind == "f" else v
            _close(rc.EmaAll32([vv], ikey, nb, fn, (rate, t, filt, reset))[0], orc.EmaAll32([vv], ikey, nb, fn, (rate, t, filt, reset))[0], f"ema {fn}")
        for fn, p in ((200, 2), (201, 3), (202, 1), (203, -1), (203, 2), (204, 1), (204, -1), (205, 3), (206, 2)):
            _close(rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [fn], [0], [nb + 1], p)[0],
                   orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [fn], [0], [nb + 1], p)[0], f"rolling {fn} {p}")
        for q, isnan in ((0.5, True), (0.25, False), (1.0, True)):
            fp = orc.quantile_func_param(q, isnan)
            _close(rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [106], [1], [nb + 1], fp)[0],
                   orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [106], [1], [nb + 1], fp)[0], f"quantile {q} {isnan}")
        _close(rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [104], [1], [nb + 1], 0)[0],
               orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [104], [1], [nb + 1], 0)[0], "mode")
    for mode in (0, 1):
        np.testing.assert_array_equal(rc.MakeiFirst(ikey, nb, filt, mode), orc.MakeiFirst(ikey, nb, filt, mode))
        np.testing.assert_array_equal(rc.MakeiNext(ikey, nb, mode), orc.MakeiNext(ikey, nb, mode))


@pytest.mark.parametrize("n", [s for s in SIZES if s > 0])
def test_fuzz_sort_and_membership(rc, n):
    rng = np.random.default_rng(3000 + n)
    a = rng.integers(-4, 4, n).astype(np.int16)
    f = np.round(rng.normal(0, 2, n), 0)
    f[rng.random(n) < 0.1] = np.nan
    s = rng.choice(np.array([b"", b"a", b"ab", b"b"], dtype="S2"), n)
    for keys, asc in (([a], True), ([f], True), ([s, a], True), ([a, f, s], True), ([f, a], [False, True])):
        got = rc.LexSort32(keys, ascending=asc)
        exp = orc.LexSort32(keys, ascending=asc)
        np.testing.assert_array_equal(got, exp)
        if asc is True:
            gik, gifk, gcnt = rc.GroupFromLexSort(exp, keys)
            eik, eifk, ecnt = orc.GroupFromLexSort(exp, keys)
            np.testing.assert_array_equal(gik, eik); np.testing.assert_array_equal(gifk, eifk); np.testing.assert_array_equal(gcnt, ecnt)
    np.testing.assert_array_equal(rc.ReverseShuffle(exp), orc.ReverseShuffle(exp))
    b = rng.integers(-2, 6, max(1, n // 4)).astype(np.int16)
    for fn in ("IsMember32", "IsMemberCategorical"):
        gf, gi = getattr(rc, fn)(a, b)
        ef, ei = getattr(orc, fn)(a, b)
        assert gi.dtype == ei.dtype
        np.testing.assert_array_equal(gf, ef); np.testing.assert_array_equal(gi, ei)
    gf, gi = rc.IsMember32(f, f[: max(1, n // 3)])
    ef, ei = orc.IsMember32(f, f[: max(1, n // 3)])
    np.testing.assert_array_equal(gf, ef); np.testing.assert_array_equal(gi, ei)
