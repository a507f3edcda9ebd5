"""GPU parity: packing, first/nth/last, cumsum, MakeiFirst/MakeiNext, LexSort32, GroupFromLexSort, ReverseShuffle
vs the CPU oracle and vs NumPy (np.lexsort / np.bincount / stable argsort), following the reference's own tests
(riptable/tests/test_lexsort.py, test_grouping.py, test_groupby.py).  Integer outputs are bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import riptide_oracle as orc


@pytest.fixture(scope="module")
def rc():
    import riptable_b200.rc as rc

    return rc


# ---------------------------------------------------------------------------------------------------- a7
def test_groupbypack_docstring(rc):
    # rt_numpy.py:1926-1933
    np.random.seed(12345)
    c = np.random.randint(0, 8, 10_000)
    ik, ifk, u = rc.MultiKeyGroupBy32([c], 0, None, 2)
    cnt, ig, ifg = rc.BinCount(ik, u + 1, pack=True)
    assert ifk.tolist() == [0, 1, 3, 4, 6, 14, 18, 20]
    assert cnt.tolist() == [0, 1213, 1252, 1296, 1226, 1252, 1283, 1275, 1203]
    assert ifg.tolist() == [0, 0, 1213, 2465, 3761, 4987, 6239, 7522, 8797]
    assert ig[:3].tolist() == [0, 9, 21] and ig[-3:].tolist() == [9988, 9991, 9992]
    assert ig.dtype == np.int32 and ifg.dtype == np.int32 and cnt.dtype == np.int32


@pytest.mark.parametrize("kdt,nb", [(np.int8, 7), (np.int8, 120), (np.int16, 300), (np.int16, 20_000), (np.int32, 70_000),
                                    (np.int32, 3_000_000), (np.int64, 257)])
@pytest.mark.parametrize("n", [1, 4097, 500_003])
def test_pack_variants(rc, kdt, nb, n):
    rng = np.random.default_rng(nb + n)
    ikey = rng.integers(0, nb, n).astype(kdt)
    cnt, ig, ifg = rc.BinCount(ikey, nb, pack=True)
    ecnt, eig, eifg = orc.BinCount(ikey, nb, pack=True)
    np.testing.assert_array_equal(cnt, ecnt)
    np.testing.assert_array_equal(ig, eig)
    np.testing.assert_array_equal(ifg, eifg)
    ig2, ifg2 = rc.GroupFromBinCount(ikey, cnt)
    np.testing.assert_array_equal(ig2, eig)
    np.testing.assert_array_equal(ifg2, eifg)
    ig3, ifg3, cnt3 = rc.GroupByPack32(ikey, None, nb)
    np.testing.assert_array_equal(ig3, eig)
    np.testing.assert_array_equal(ifg3, eifg)
    np.testing.assert_array_equal(cnt3, ecnt)


def test_pack_skewed_and_constant(rc):
    n = 300_000
    for ikey in (np.zeros(n, np.int32), np.full(n, 5, np.int32), np.repeat(np.arange(300, dtype=np.int32), 1000),
                 (np.arange(n) % 2).astype(np.int8) * 77):
        nb = int(ikey.max()) + 1
        cnt, ig, ifg = rc.BinCount(ikey, nb, pack=True)
        ecnt, eig, eifg = orc.BinCount(ikey, nb, pack=True)
        np.testing.assert_array_equal(ig, eig)
        np.testing.assert_array_equal(ifg, eifg)
        np.testing.assert_array_equal(cnt, ecnt)
    with pytest.raises(ValueError):
        rc.BinCount(np.array([0, 1, 9], np.int32), 3, pack=True)


# ---------------------------------------------------------------------------------------------------- a8
@pytest.mark.parametrize("vdt", [np.bool_, np.int8, np.uint16, np.int32, np.uint64, np.float32, np.float64, "S5", "U3", "S16"])
def test_first_nth_last(rc, vdt):
    rng = np.random.default_rng(3)
    n, nb = 100_000, 3000
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    ikey[ikey == 7] = 8  # an empty bin
    vdt = np.dtype(vdt)
    if vdt.kind in "SU":
        pool = np.array(["", "a", "bc", "zzzzz", "hello world pad16"], dtype=vdt)
        v = pool[rng.integers(0, len(pool), n)]
    elif vdt.kind == "b":
        v = rng.integers(0, 2, n).astype(vdt)
    else:
        v = (rng.random(n) * 100).astype(vdt)
    cnt, ig, ifg = rc.BinCount(ikey, nb + 1, pack=True)
    for fn, param in [(100, 0), (102, 0), (101, 0), (101, 2), (101, -1), (101, -3), (101, 1000)]:
        for lo in (1, 0):
            got = rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [fn], [lo], [nb + 1], param)[0]
            exp = orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [fn], [lo], [nb + 1], param)[0]
            assert got.dtype == exp.dtype
            if vdt.kind == "f":
                np.testing.assert_array_equal(got[lo:].view(f"u{vdt.itemsize}"), exp[lo:].view(f"u{vdt.itemsize}"))
            else:
                np.testing.assert_array_equal(got[lo:], exp[lo:])


def test_nth_kat(rc):
    # tests/test_groupby.py:137-159 style: nth past the end -> invalid, negative counts from the end
    ikey = np.array([1, 1, 2, 1, 2, 3], dtype=np.int8)
    v = np.array([10.0, 11.0, 20.0, 12.0, 21.0, 30.0])
    cnt, ig, ifg = rc.BinCount(ikey, 4, pack=True)
    first = rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, 3, [100], [1], [4], 0)[0]
    last = rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, 3, [102], [1], [4], 0)[0]
    nth1 = rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, 3, [101], [1], [4], 1)[0]
    assert first[1:].tolist() == [10.0, 20.0, 30.0] and last[1:].tolist() == [12.0, 21.0, 30.0]
    assert nth1[1:3].tolist() == [11.0, 21.0] and np.isnan(nth1[3])


# ---------------------------------------------------------------------------------------------------- a9
def _cmp_cumsum(got, exp, v):
    assert got.dtype == exp.dtype
    if exp.dtype.kind == "f":
        scale = float(np.nansum(np.abs(v.astype(np.float64)))) + 1.0
        rtol = 1e-5 if exp.dtype == np.float32 else 1e-12
        np.testing.assert_allclose(got, exp, rtol=rtol, atol=rtol * scale, equal_nan=True)
    else:
        np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("vdt", [np.bool_, np.int8, np.int32, np.int64, np.uint64, np.float32, np.float64])
@pytest.mark.parametrize("nb", [1, 40, 25_000])
def test_cumsum(rc, vdt, nb):
    rng = np.random.default_rng(nb)
    n = 200_003
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    if np.dtype(vdt).kind == "f":
        v = (rng.random(n) * 2 - 1).astype(vdt)
    elif np.dtype(vdt).kind == "b":
        v = rng.integers(0, 2, n).astype(vdt)
    else:
        v = rng.integers(0, 100, n).astype(vdt)
    filt = rng.random(n) < 0.7
    reset = rng.random(n) < 0.01
    for fp in [(0.0, None, None, None), (0.0, None, filt, None), (0.0, None, filt, reset), (0.0, None, None, reset)]:
        got = rc.EmaAll32([v], ikey, nb, 300, fp)[0]
        exp = orc.EmaAll32([v], ikey, nb, 300, fp)[0]
        _cmp_cumsum(got, exp, v)


@pytest.fixture
def blocked_cumsum(monkeypatch):
    """RTB_CUMSUM_BLOCKED=2: the sort-free (blocked) form of GB_CUMSUM runs whenever it can, however small the call."""
    monkeypatch.setenv("RTB_CUMSUM_BLOCKED", "2")


@pytest.mark.parametrize("vdt", [np.bool_, np.int8, np.uint16, np.int32, np.int64, np.uint64, np.float32, np.float64])
@pytest.mark.parametrize("nb,kdt", [(1, np.int8), (7, np.int16), (40, np.int32), (3000, np.int64), (3500, np.int32), (16_000, np.int32)])
def test_cumsum_blocked(rc, blocked_cumsum, vdt, nb, kdt):
    """The blocked form (csrc/cumsum_blocked.cu) against the oracle: few bins (a chunk per warp, whole warps of one bin),
    more bins (a chunk per CTA, a table slot per bin of the tile), every iKey width, the op filter, bin 0, ragged last
    tile / chunk."""
    rng = np.random.default_rng(nb + 17)
    n = 300_017
    ikey = rng.integers(0, nb + 1, n).astype(kdt)
    if np.dtype(vdt).kind == "f":
        v = (rng.random(n) * 2 - 1).astype(vdt)
    elif np.dtype(vdt).kind == "b":
        v = rng.integers(0, 2, n).astype(vdt)
    else:
        v = rng.integers(0, 100, n).astype(vdt)
    filt = rng.random(n) < 0.7
    for fp in [(0.0, None, None, None), (0.0, None, filt, None)]:
        got = rc.EmaAll32([v], ikey, nb, 300, fp)[0]
        exp = orc.EmaAll32([v], ikey, nb, 300, fp)[0]
        _cmp_cumsum(got, exp, v)


def test_cumsum_blocked_bad_ikey(rc, blocked_cumsum):
    """An iKey outside [0, unique_rows] is an error on the blocked form as on the packed one."""
    for nb, bad in ((2, 9), (5000, 5001), (2, -1)):
        ikey = np.array([1, 2, bad, 1], dtype=np.int32)
        with pytest.raises(ValueError):
            rc.EmaAll32([np.arange(4.0)], ikey, nb, 300, (0.0, None, None, None))


def test_cumsum_blocked_special_values(rc, blocked_cumsum):
    """NaN / Inf poison the rest of their group exactly as a sequential sum does, across tile and chunk boundaries;
    negative integers and the int64 sentinel are ordinary values (tests/test_merge.py:6425-6426)."""
    rng = np.random.default_rng(3)
    n = 150_000
    ikey = rng.integers(0, 51, n).astype(np.int32)
    v = rng.normal(0, 1, n)
    v[rng.integers(0, n, 6)] = np.nan
    v[rng.integers(0, n, 4)] = np.inf
    v[rng.integers(0, n, 4)] = -np.inf
    got = rc.EmaAll32([v], ikey, 50, 300, (0.0, None, None, None))[0]
    exp = orc.EmaAll32([v], ikey, 50, 300, (0.0, None, None, None))[0]
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    np.testing.assert_array_equal(np.isposinf(got), np.isposinf(exp))
    np.testing.assert_array_equal(np.isneginf(got), np.isneginf(exp))
    fin = np.isfinite(exp)
    np.testing.assert_allclose(got[fin], exp[fin], rtol=1e-12, atol=1e-9)
    vi = rng.integers(-10**15, 10**15, n).astype(np.int64)
    vi[::5000] = np.iinfo(np.int64).min
    got = rc.EmaAll32([vi], ikey, 50, 300, (0.0, None, None, None))[0]
    exp = orc.EmaAll32([vi], ikey, 50, 300, (0.0, None, None, None))[0]
    np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("nb", [500, 12_000])
def test_cumsum_blocked_matches_sorted_form_at_size(rc, monkeypatch, nb):
    """8 Mi rows, 500 bins (a chunk per warp) / 12 000 bins (a chunk per CTA): the call takes the blocked form on its own
    (no override); integers must equal the sort-based form and the oracle bit for bit, floats within the summation
    tolerance."""
    rng = np.random.default_rng(8)
    n = 8 << 20
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    vi = rng.integers(-1000, 1000, n).astype(np.int32)
    vf = rng.random(n)
    monkeypatch.setenv("RTB_CUMSUM_BLOCKED", "0")
    si = rc.EmaAll32([vi], ikey, nb, 300, (0.0, None, None, None))[0]
    sf = rc.EmaAll32([vf], ikey, nb, 300, (0.0, None, None, None))[0]
    monkeypatch.delenv("RTB_CUMSUM_BLOCKED")
    bi = rc.EmaAll32([vi], ikey, nb, 300, (0.0, None, None, None))[0]
    bf = rc.EmaAll32([vf], ikey, nb, 300, (0.0, None, None, None))[0]
    np.testing.assert_array_equal(bi, si)
    np.testing.assert_allclose(bf, sf, rtol=1e-12, atol=1e-7, equal_nan=True)
    np.testing.assert_array_equal(bi, orc.EmaAll32([vi], ikey, nb, 300, (0.0, None, None, None))[0])


def test_cumsum_kat(rc):
    # tests/test_groupby.py:555-602: op-filter False -> carry the running value; bin 0 -> invalid
    ikey = np.array([1, 2, 1, 0, 2, 1], dtype=np.int8)
    v = np.array([1, 10, 2, 100, 20, 3], dtype=np.int32)
    out = rc.EmaAll32([v], ikey, 2, 300, (0.0, None, None, None))[0]
    assert out.dtype == np.int64
    assert out.tolist() == [1, 10, 3, np.iinfo(np.int64).min, 30, 6]
    f = np.array([True, True, False, True, True, True])
    out = rc.EmaAll32([v], ikey, 2, 300, (0.0, None, f, None))[0]
    assert out.tolist() == [1, 10, 1, np.iinfo(np.int64).min, 30, 4]
    s = np.array([b"a"] * 6)
    assert rc.EmaAll32([s], ikey, 2, 300, (0.0, None, None, None))[0] is None


# ---------------------------------------------------------------------------------------------------- 8f.2
_ALL_NUM = [np.bool_, np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64]


@pytest.mark.parametrize("vdt", _ALL_NUM)
@pytest.mark.parametrize("nb,bad_frac", [(1, 0.0), (40, 0.02), (25_000, 0.3), (40, 1.0)])
def test_cum_min_max(rc, vdt, nb, bad_frac):
    rng = np.random.default_rng(nb * 7 + int(bad_frac * 100))
    n = 150_001
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    dt = np.dtype(vdt)
    if dt.kind == "f":
        v = rng.normal(0, 1e4, n).astype(vdt)
        v[rng.random(n) < bad_frac / 2] = np.inf
        v[rng.random(n) < bad_frac / 2] = -np.inf
    elif dt.kind == "b":
        v = rng.integers(0, 2, n).astype(vdt)
    else:
        info = np.iinfo(dt)
        v = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    if dt.kind != "b":
        v[rng.random(n) < bad_frac] = orc.invalid_for(dt)
    filt = rng.random(n) < 0.7
    reset = rng.random(n) < 0.01
    for func in (306, 307, 308, 309):
        for fp in [(0.0, None, None, None), (0.0, None, filt, reset)]:
            got = rc.EmaAll32([v], ikey, nb, func, fp)[0]
            exp = orc.EmaAll32([v], ikey, nb, func, fp)[0]
            assert got.dtype == exp.dtype == dt
            np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("vdt", _ALL_NUM)
@pytest.mark.parametrize("nb,bad_frac", [(1, 0.0), (40, 0.02), (3500, 0.3), (16_000, 0.05), (40, 1.0)])
def test_cum_min_max_blocked(rc, blocked_cumsum, vdt, nb, bad_frac):
    """Running minima / maxima on the sort-free blocked form (csrc/cumsum_blocked.cu): the order-preserving codes with
    "nothing yet" as the identity and "poisoned" as the absorbing extreme; skipna and plain variants, every dtype,
    chunk-per-warp and chunk-per-CTA bin counts, a group that is invalid throughout."""
    rng = np.random.default_rng(nb * 7 + int(bad_frac * 100) + 1)
    n = 200_003
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    dt = np.dtype(vdt)
    if dt.kind == "f":
        v = rng.normal(0, 1e4, n).astype(vdt)
        v[rng.random(n) < bad_frac / 2] = np.inf
        v[rng.random(n) < bad_frac / 2] = -np.inf
    elif dt.kind == "b":
        v = rng.integers(0, 2, n).astype(vdt)
    else:
        info = np.iinfo(dt)
        v = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    if dt.kind != "b":
        v[rng.random(n) < bad_frac] = orc.invalid_for(dt)
    filt = rng.random(n) < 0.7
    for func in (306, 307, 308, 309):
        for fp in [(0.0, None, None, None), (0.0, None, filt, None)]:
            got = rc.EmaAll32([v], ikey, nb, func, fp)[0]
            exp = orc.EmaAll32([v], ikey, nb, func, fp)[0]
            assert got.dtype == exp.dtype == dt
            np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("vdt", [np.bool_, np.int8, np.int32, np.int64, np.uint64, np.float32, np.float64])
@pytest.mark.parametrize("nb", [1, 40, 25_000])
def test_cumprod(rc, vdt, nb):
    rng = np.random.default_rng(nb + 5)
    n = 120_007
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    if np.dtype(vdt).kind == "f":
        v = np.exp(rng.normal(0, 0.01, n)).astype(vdt)  # running products stay in range
    elif np.dtype(vdt).kind == "b":
        v = (rng.random(n) < 0.999).astype(vdt)
    else:
        v = rng.integers(1, 4, n).astype(vdt)  # int64 products wrap; the wrap is part of the contract
    filt = rng.random(n) < 0.7
    reset = rng.random(n) < 0.05
    for fp in [(0.0, None, None, None), (0.0, None, filt, reset)]:
        got = rc.EmaAll32([v], ikey, nb, 302, fp)[0]
        exp = orc.EmaAll32([v], ikey, nb, 302, fp)[0]
        assert got.dtype == exp.dtype
        if exp.dtype.kind == "f":
            # the scan multiplies in tree order: a few ulp per 2x of segment length
            np.testing.assert_allclose(got, exp, rtol=2e-5 if exp.dtype == np.float32 else 1e-11, equal_nan=True)
        else:
            np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("vdt", _ALL_NUM + ["S5", "U3"])
@pytest.mark.parametrize("nb", [1, 37, 20_000])
def test_rolling_ops(rc, vdt, nb):
    rng = np.random.default_rng(nb + 11)
    n = 90_001
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)  # bin 0 present: it is rolled as a group of its own
    dt = np.dtype(vdt)
    if dt.kind == "f":
        v = rng.normal(0, 100, n).astype(dt)
        v[rng.random(n) < 0.1] = np.nan
    elif dt.kind == "b":
        v = rng.integers(0, 2, n).astype(dt)
    elif dt.kind in "iu":
        info = np.iinfo(dt)
        v = rng.integers(max(info.min, -1000), min(info.max, 1000), n).astype(dt)
        v[rng.random(n) < 0.1] = orc.invalid_for(dt)
    else:
        v = rng.choice(np.array(["", "a", "bc", "xyz", "hello"]), n).astype(dt)
    cnt, ig, ifg = orc.BinCount(ikey, nb + 1, pack=True)
    cases = [(203, -2), (203, 3), (203, 0), (202, 1), (202, -3), (200, 3), (201, 1), (201, 40), (205, 5), (206, 4), (204, 1), (204, -1)]
    for fn, p in cases:
        exp = orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [fn], [0], [nb + 1], p)[0]
        got = rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [fn], [0], [nb + 1], p)[0]
        if exp is None:
            assert got is None, (fn, vdt)
            continue
        assert got.dtype == exp.dtype, (fn, vdt)
        if exp.dtype == np.float32 and fn in (200, 201):
            np.testing.assert_allclose(got, exp, rtol=1e-6, equal_nan=True)  # f64 accumulation, one rounding to f32
        else:
            np.testing.assert_array_equal(got, exp, err_msg=f"func {fn} param {p} dtype {vdt}")


@pytest.mark.parametrize("vdt", _ALL_NUM)
@pytest.mark.parametrize("nb", [1, 37, 20_000])
def test_quantile(rc, vdt, nb):
    rng = np.random.default_rng(nb + 3)
    n = 120_001
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    dt = np.dtype(vdt)
    if dt.kind == "f":
        v = rng.normal(0, 1e4, n).astype(dt)
        v[rng.random(n) < 0.05] = np.inf
        v[rng.random(n) < 0.05] = -np.inf
        v[rng.random(n) < 0.1] = np.nan
    elif dt.kind == "b":
        v = rng.integers(0, 2, n).astype(dt)
    else:
        info = np.iinfo(dt)
        v = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
        v[rng.random(n) < 0.1] = orc.invalid_for(dt)
    cnt, ig, ifg = orc.BinCount(ikey, nb + 1, pack=True)
    for q, is_nan in ((0.5, True), (0.5, False), (0.0, True), (1.0, True), (0.29, True), (0.999, False)):
        fp = orc.quantile_func_param(q, is_nan)
        for lo in (0, 1):
            exp = orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [106], [lo], [nb + 1], fp)[0]
            got = rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [106], [lo], [nb + 1], fp)[0]
            assert got.dtype == np.float64
            np.testing.assert_array_equal(got, exp, err_msg=f"q {q} nan {is_nan} dtype {vdt}")
    np.testing.assert_array_equal(rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [103], [1], [nb + 1], 0)[0],
                                  orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [103], [1], [nb + 1], 0)[0])
    s = np.array([b"a"] * n)
    assert rc.GroupByAllPack32([s], ikey, ig, ifg, cnt, nb, [103], [1], [nb + 1], 0)[0] is None


@pytest.mark.parametrize("vdt", [np.bool_, np.int8, np.int32, np.uint64, np.float32, np.float64])
@pytest.mark.parametrize("nb", [1, 9, 4000])
def test_ema(rc, vdt, nb):
    rng = np.random.default_rng(nb + 17)
    n = 30_011  # the oracle walks the rows in Python
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    dt = np.dtype(vdt)
    v = (rng.normal(0, 10, n).astype(dt) if dt.kind == "f" else rng.integers(0, 2 if dt.kind == "b" else 100, n).astype(dt))
    filt = rng.random(n) < 0.8
    reset = rng.random(n) < 0.01
    for tdt in (np.float64, np.int64, np.int32):
        t = np.cumsum(rng.integers(0, 5, n)).astype(tdt)
        for fn, rate in ((301, 0.05), (304, 0.3), (305, 0.9)):
            for fp in [(rate, t, None, None), (rate, t, filt, reset)]:
                exp = orc.EmaAll32([v], ikey, nb, fn, fp)[0]
                got = rc.EmaAll32([v], ikey, nb, fn, fp)[0]
                assert got.dtype == exp.dtype
                tol = 2e-5 if exp.dtype == np.float32 else 1e-11
                scale = float(np.nanmax(np.abs(exp))) + 1.0
                np.testing.assert_allclose(got, exp, rtol=tol, atol=tol * scale, equal_nan=True, err_msg=f"func {fn} time {tdt}")
    assert rc.EmaAll32([np.array([b"a"] * n)], ikey, nb, 305, (0.5, None, None, None))[0] is None
    with pytest.raises(ValueError):
        rc.EmaAll32([v], ikey, nb, 301, (0.5, None, None, None))


def test_ema_nonfinite_values_stay_inside_their_group(rc):
    """a NaN / Inf value (or an overflowing decay weight on unsorted times) in one bin must not reach the
    bins that follow it in packed order, and reset_filter restarts a group after a NaN."""
    rng = np.random.default_rng(77)
    n, nb = 20_000, 50
    ikey = rng.integers(1, nb + 1, n).astype(np.int32)
    t = np.cumsum(rng.integers(0, 4, n)).astype(np.float64)
    for bad in (np.nan, np.inf, -np.inf):
        v = rng.normal(0, 1, n)
        rows1 = np.flatnonzero(ikey == 1)
        v[rows1[3]] = bad  # early in bin 1, the first bin of the packed order
        v[np.flatnonzero(ikey == 7)[0]] = bad  # and as the very first value of a group
        reset = np.zeros(n, dtype=bool)
        reset[rows1[10]] = True  # bin 1 restarts after its bad value
        for fn, rate in ((301, 0.05), (304, 0.3), (305, 0.9)):
            for fp in [(rate, t, None, None), (rate, t, None, reset)]:
                exp = orc.EmaAll32([v], ikey, nb, fn, fp)[0]
                got = rc.EmaAll32([v], ikey, nb, fn, fp)[0]
                other = (ikey != 1) & (ikey != 7)
                assert np.isfinite(exp[other]).all() and np.isfinite(got[other]).all(), f"func {fn}: leaked out of its bin"
                # rows of the two affected groups that follow the bad value (up to a reset) are non-finite on both sides;
                # WHICH non-finite value is not compared: the row-by-row recurrence keeps Inf * w = Inf for ever, while a
                # scan composes the weights first and their product underflows to 0 over a long run (0 * Inf = NaN)
                tainted = np.zeros(n, dtype=bool)
                rows7 = np.flatnonzero(ikey == 7)
                tainted[rows1[3:10] if fp[3] is not None else rows1[3:]] = True
                tainted[rows7] = True
                assert (~np.isfinite(exp[tainted])).all() and (~np.isfinite(got[tainted])).all()
                np.testing.assert_allclose(got[~tainted], exp[~tainted], rtol=1e-11, atol=1e-9, err_msg=f"func {fn} bad {bad}")
                if fp[3] is not None:
                    assert np.isfinite(got[rows1[10:]]).all(), "reset_filter did not clear the group"
    # unsorted times: exp(-rate * dt) overflows to Inf inside bin 2 only
    v = rng.normal(0, 1, n)
    t2 = t.copy()
    rows2 = np.flatnonzero(ikey == 2)
    t2[rows2[5]] = -1e6
    for fn in (301, 304):
        exp = orc.EmaAll32([v], ikey, nb, fn, (1.0, t2, None, None))[0]
        got = rc.EmaAll32([v], ikey, nb, fn, (1.0, t2, None, None))[0]
        assert np.isfinite(got[ikey != 2]).all()
        np.testing.assert_allclose(got[ikey != 2], exp[ikey != 2], rtol=1e-11, atol=1e-9)


def test_ema_docstring_kats(rc):
    test = np.arange(10)
    g = (np.arange(10) % 3 + 1).astype(np.int8)
    normal = rc.EmaAll32([test], g, 3, 304, (1.0, np.arange(10), None, None))[0]
    weighted = rc.EmaAll32([test], g, 3, 305, (0.5, None, None, None))[0]
    assert np.round(normal, 2).tolist() == [0.00, 1.00, 2.00, 2.85, 3.85, 4.85, 5.84, 6.84, 7.84, 8.84]
    assert np.round(weighted, 2).tolist() == [0.00, 1.00, 2.00, 1.50, 2.50, 3.50, 3.75, 4.75, 5.75, 6.38]


@pytest.mark.parametrize("vdt", [np.bool_, np.int8, np.uint16, np.int32, np.int64, np.uint64, np.float32, np.float64])
@pytest.mark.parametrize("nb", [1, 37, 5000])
def test_mode(rc, vdt, nb):
    rng = np.random.default_rng(nb + 29)
    n = 80_001
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    dt = np.dtype(vdt)
    v = rng.integers(0, 2 if dt.kind == "b" else 12, n).astype(dt)
    if dt.kind != "b":
        v[rng.random(n) < 0.2] = orc.invalid_for(dt)
    cnt, ig, ifg = orc.BinCount(ikey, nb + 1, pack=True)
    for lo in (0, 1):
        exp = orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [104], [lo], [nb + 1], 0)[0]
        got = rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [104], [lo], [nb + 1], 0)[0]
        assert got.dtype == exp.dtype == dt
        np.testing.assert_array_equal(got, exp)
    assert rc.GroupByAllPack32([np.array([b"a"] * n)], ikey, ig, ifg, cnt, nb, [104], [1], [nb + 1], 0)[0] is None


@pytest.mark.parametrize("vdt", [np.float64, np.float32, np.int32, np.uint8, np.bool_])
@pytest.mark.parametrize("nb,window", [(1, 150), (37, 3), (37, 40), (300, 900), (5000, 1)])
def test_rolling_quantile(rc, vdt, nb, window):
    rng = np.random.default_rng(nb + window)
    n = 20_011
    ikey = rng.integers(0, nb + 1, n).astype(np.int32)
    dt = np.dtype(vdt)
    if dt.kind == "f":
        v = rng.normal(0, 100, n).astype(dt)
        v[rng.random(n) < 0.05] = np.inf
        v[rng.random(n) < 0.05] = -np.inf
        v[rng.random(n) < 0.15] = np.nan
    elif dt.kind == "b":
        v = rng.integers(0, 2, n).astype(dt)
    else:
        v = rng.integers(0, 200, n).astype(dt)
        v[rng.random(n) < 0.15] = orc.invalid_for(dt)
    cnt, ig, ifg = orc.BinCount(ikey, nb + 1, pack=True)
    for q in (0.5, 0.0, 1.0, 0.29):
        fp = orc.rolling_quantile_func_param(q, window)
        exp = orc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [207], [0], [nb + 1], fp)[0]
        got = rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [207], [0], [nb + 1], fp)[0]
        assert got.dtype == np.float64
        np.testing.assert_array_equal(got, exp, err_msg=f"q {q} window {window} dtype {vdt}")
    with pytest.raises(ValueError):
        rc.GroupByAllPack32([v], ikey, ig, ifg, cnt, nb, [207], [0], [nb + 1], orc.rolling_quantile_func_param(0.5, 5000))


def test_rolling_kat(rc):
    # tests/test_fastarray.py:1297-1305 through a single group
    x = np.array([1, 2, 3, np.nan, np.nan, np.nan])
    one = np.ones(6, np.int8)
    c1, g1, f1 = orc.BinCount(one, 2, pack=True)
    assert rc.GroupByAllPack32([x], one, g1, f1, c1, 1, [201], [0], [2], 3)[0].tolist() == [1, 3, 6, 5, 3, 0]
    with pytest.raises(ValueError):
        rc.GroupByAllPack32([x], one, g1, f1, c1, 1, [200], [0], [2], 0)


# ---------------------------------------------------------------------------------------------------- a10
def test_makeifirst_ilast_inext(rc):
    rng = np.random.default_rng(8)
    for kdt, nb, n in [(np.int8, 3, 524_288), (np.int16, 1000, 100_001), (np.int32, 50_000, 300_000), (np.int64, 9, 77)]:
        key = rng.integers(0, nb + 1, n).astype(kdt)
        key[key == 2] = 1
        f = rng.random(n) < 0.5
        for mode in (0, 1):
            np.testing.assert_array_equal(rc.MakeiFirst(key, nb, None, mode), orc.MakeiFirst(key, nb, None, mode))
            np.testing.assert_array_equal(rc.MakeiFirst(key, nb, f, mode), orc.MakeiFirst(key, nb, f, mode))
            np.testing.assert_array_equal(rc.MakeiFirst(key, nb + 1, f, mode), orc.MakeiFirst(key, nb + 1, f, mode))
            np.testing.assert_array_equal(rc.MakeiNext(key, nb, mode), orc.MakeiNext(key, nb, mode))
    # tests/test_grouping.py:210-296 shape: int8 keys on 2**19 rows, the last row index does not fit the key dtype
    key = np.tile(np.array([1, 2, 3], dtype=np.int8), 174_763)[:524_288]
    ilast = rc.MakeiFirst(key, 3, None, 1)
    assert ilast.dtype == np.int32 and int(ilast[1:].max()) == 524_287


def test_reverse_shuffle(rc):
    rng = np.random.default_rng(2)
    p = rng.permutation(1_000_003).astype(np.int32)
    np.testing.assert_array_equal(rc.ReverseShuffle(p), orc.ReverseShuffle(p))


# ---------------------------------------------------------------------------------------------------- a11
def _lex_keys(rng, n):
    f = rng.random(n) * 10 - 5
    f[rng.random(n) < 0.05] = np.nan
    f[rng.random(n) < 0.02] = 0.0
    f[rng.random(n) < 0.02] = -0.0
    f[rng.random(n) < 0.01] = np.inf
    f[rng.random(n) < 0.01] = -np.inf
    return {
        "i8": rng.integers(-128, 128, n).astype(np.int8),
        "u8": rng.integers(0, 256, n).astype(np.uint8),
        "i16": rng.integers(-300, 300, n).astype(np.int16),
        "u16": rng.integers(0, 70, n).astype(np.uint16),
        "i32": rng.integers(-(2**31), 2**31 - 1, n).astype(np.int32),
        "u32": rng.integers(0, 2**32 - 1, n).astype(np.uint32),
        "i64": rng.integers(-(2**62), 2**62, n).astype(np.int64),
        "u64": rng.integers(0, 2**63, n).astype(np.uint64) * np.uint64(2),
        "f64": f,
        "f32": (f * 3).astype(np.float32),
        "b": rng.integers(0, 2, n).astype(np.bool_),
        "S7": rng.choice(np.array([b"", b"a", b"ab", b"abc\x00d", b"zzzzzzz", b"B"], dtype="S7"), n),
        "S20": rng.choice(np.array([b"hello world 1234567", b"hello world 1234568", b"hello", b"hello world"], dtype="S20"), n),
        "U5": rng.choice(np.array(["", "a", "ab", "été", "zzzzz", "\U0001F600x"], dtype="U5"), n),
    }


@pytest.mark.parametrize("n", [1, 2, 1000, 250_007])
def test_lexsort_single_key_every_dtype(rc, n):
    rng = np.random.default_rng(n)
    for name, k in _lex_keys(rng, n).items():
        got = rc.LexSort32([k])
        assert got.dtype == np.int32
        np.testing.assert_array_equal(got, np.lexsort([k]), err_msg=name)


def test_lexsort_multikey_matches_numpy(rc):
    # tests/test_lexsort.py:21-103: mixed keys, last key primary
    rng = np.random.default_rng(77)
    n = 1_000_000
    ks = _lex_keys(rng, n)
    low = rng.integers(0, 5, n).astype(np.int32)
    for combo in (["i32", "u16"], ["f64", "i16", "b"], ["S7", "u8"], ["i64", "f32", "S20", "U5", "i8"], ["u64", "u32"]):
        keys = [ks[c] for c in combo] + [low]
        np.testing.assert_array_equal(rc.LexSort32(keys), np.lexsort(keys), err_msg=str(combo))
    np.testing.assert_array_equal(rc.LexSort64([ks["i16"], low]), np.lexsort([ks["i16"], low]))


def test_lexsort_index_and_ascending(rc):
    rng = np.random.default_rng(5)
    n = 200_000
    a = rng.integers(0, 50, n).astype(np.int32)
    f = rng.random(n)
    f[rng.random(n) < 0.1] = np.nan
    s = rng.choice(np.array([b"x", b"yy", b"", b"zzz"], dtype="S3"), n)
    idx = np.flatnonzero(rng.random(n) < 0.4).astype(np.int32)
    np.testing.assert_array_equal(rc.LexSort32([f, a], index=idx), orc.LexSort32([f, a], index=idx))
    for asc in (False, [True, False], [False, True], [False, False]):
        np.testing.assert_array_equal(rc.LexSort32([f, a], ascending=asc), orc.LexSort32([f, a], ascending=asc))
    np.testing.assert_array_equal(rc.LexSort32([s, a], ascending=[False, True]), orc.LexSort32([s, a], ascending=[False, True]))
    np.testing.assert_array_equal(rc.LexSort32([f], index=idx, ascending=False), orc.LexSort32([f], index=idx, ascending=False))
    assert len(rc.LexSort32([np.zeros(0)])) == 0
    with pytest.raises(ValueError):
        rc.LexSort32([np.zeros(3), np.zeros(4)])


def test_lexsort_record_array(rc):
    rng = np.random.default_rng(1)
    n = 50_000
    a, b = rng.integers(0, 9, n).astype(np.int64), rng.integers(0, 9, n).astype(np.int16)
    rec = np.rec.fromarrays([a, b], names=["a", "b"])
    # a record array is sorted "lumped together" (rt_numpy.py:2110-2112, 2176-2178): bytewise over the whole record
    raw = np.ascontiguousarray(rec).view(np.uint8).reshape(n, rec.dtype.itemsize)
    exp = np.lexsort(tuple(raw[:, j] for j in range(rec.dtype.itemsize - 1, -1, -1)))
    np.testing.assert_array_equal(rc.LexSort32(rec), exp)
    ik, ifk, cnt = rc.GroupFromLexSort(exp.astype(np.int32), rec)
    eik, eifk, ecnt = orc.GroupFromLexSort(exp.astype(np.int32), [a, b])
    np.testing.assert_array_equal(ik, eik)
    np.testing.assert_array_equal(cnt, ecnt)


# ---------------------------------------------------------------------------------------------------- a12
def test_groupbylex_docstring(rc):
    # rt_numpy.py:2126-2156, replayed through riptable's own Python glue (tests/rt_glue.py)
    import rt_glue

    a = np.arange(100).astype("S")
    f = (np.arange(100) % 3).astype(bool)
    got = rt_glue.groupbylex(rc, [a], filter=f)
    exp = rt_glue.groupbylex(orc, [a], filter=f)
    for k in ("iKey", "iFirstKey", "iGroup", "iFirstGroup", "nCountGroup"):
        np.testing.assert_array_equal(got[k], exp[k], err_msg=k)
    assert got["unique_count"] == 66 and got["iFirstGroup"][0] == 66 and got["nCountGroup"][0] == 34
    assert got["iKey"][:9].tolist() == [0, 1, 9, 0, 23, 31, 0, 45, 53]
    assert got["iGroup"][:9].tolist() == [1, 10, 11, 13, 14, 16, 17, 19, 2]
    # hash == lex when keys first appear in sorted order (tests/test_lexsort.py:105-214)
    rng = np.random.default_rng(0)
    vn = rng.integers(0, 3, 10_000)
    vs = rng.choice(np.array(["a", "b", "c"]), 10_000)
    vn[:9] = [0, 0, 0, 1, 1, 1, 2, 2, 2]
    vs[:9] = ["a", "b", "c"] * 3
    gh, gl = rt_glue.groupbyhash(rc, [vn, vs], pack=True), rt_glue.groupbylex(rc, [vn, vs])
    for k in ("iKey", "iFirstKey", "unique_count", "iGroup", "iFirstGroup", "nCountGroup"):
        assert np.array_equal(gh[k], gl[k]), k
    gr = rt_glue.groupbylex(rc, [vn, vs], rec=True)
    assert gr["unique_count"] == 9 and np.array_equal(np.sort(gr["iFirstKey"]), np.sort(gl["iFirstKey"]))


@pytest.mark.parametrize("base_index", [0, 1])
def test_group_from_lexsort_vs_oracle_and_hash(rc, base_index):
    # tests/test_lexsort.py:105-256: lex grouping of sorted-first-appearance keys equals hash grouping
    rng = np.random.default_rng(4)
    n = 300_000
    f = np.round(rng.random(n) * 50)
    f[rng.random(n) < 0.05] = np.nan
    s = rng.choice(np.array([b"aa", b"b", b"", b"cccc"], dtype="S4"), n)
    i = rng.integers(-4, 4, n).astype(np.int8)
    for vals in ([f], [s], [i, s], [f, i, s]):
        sort_keys = vals[::-1]
        ig = rc.LexSort32(sort_keys)
        ik, ifk, cnt = rc.GroupFromLexSort(ig, vals, base_index=base_index)
        eik, eifk, ecnt = orc.GroupFromLexSort(np.lexsort(sort_keys).astype(np.int32), vals, base_index=base_index)
        np.testing.assert_array_equal(ik, eik)
        np.testing.assert_array_equal(ifk, eifk)
        np.testing.assert_array_equal(cnt, ecnt)
    # filter -> index
    idx = np.flatnonzero(rng.random(n) < 0.5).astype(np.int32)
    ig = rc.LexSort32([i], index=idx)
    ik, ifk, cnt = rc.GroupFromLexSort(ig, [i], base_index=base_index)
    eik, eifk, ecnt = orc.GroupFromLexSort(orc.LexSort32([i], index=idx), [i], base_index=base_index)
    np.testing.assert_array_equal(ik, eik)
    np.testing.assert_array_equal(ifk, eifk)
    np.testing.assert_array_equal(cnt, ecnt)
    # keys that first appear in sorted order: hash bins == lex bins (tests/test_lexsort.py:105-148)
    k = np.sort(rng.integers(0, 1000, n)).astype(np.int32)
    ig = rc.LexSort32([k])
    ik, ifk, cnt = rc.GroupFromLexSort(ig, [k], base_index=1)
    hk, hfk, u = rc.MultiKeyGroupBy32([k], 0, None, 2)
    np.testing.assert_array_equal(ik, hk)
    np.testing.assert_array_equal(ifk, hfk)
    np.testing.assert_array_equal(cnt, rc.BinCount(hk, u + 1))


def test_device_mode_pack_sort_chain(rc):
    import torch

    rng = np.random.default_rng(10)
    n = 400_000
    k = rng.integers(0, 3000, n).astype(np.int64)
    v = rng.random(n)
    kt, vt = torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda()
    ik, ifk, u = rc.MultiKeyGroupBy32([kt], 0, None, 2)
    cnt, ig, ifg = rc.BinCount(ik, u + 1, pack=True)
    assert ig.is_cuda and cnt.is_cuda
    first = rc.GroupByAllPack32([vt], ik, ig, ifg, cnt, u, [100], [1], [u + 1], 0)[0]
    cs = rc.EmaAll32([vt], ik, u, 300, (0.0, None, None, None))[0]
    perm = rc.LexSort32([vt, kt])
    ok, ofk, ou = orc.MultiKeyGroupBy32([k])
    np.testing.assert_array_equal(first.cpu().numpy()[1:], v[ofk])
    np.testing.assert_array_equal(perm.cpu().numpy(), np.lexsort([v, k]))
    exp = orc.EmaAll32([v], ok, ou, 300, (0.0, None, None, None))[0]
    np.testing.assert_allclose(cs.cpu().numpy(), exp, rtol=1e-12, atol=1e-9)


def test_lex_partition_matches_sorted_rank(rc):
    """rtb_lex_partition (the routing step of the multi-GPU sample sort): destination = number of splitters at or
    before the row under LexSort32's order with the global row id as the last tie-break."""
    import torch

    rng = np.random.default_rng(12)
    n = 60_000
    ks = _lex_keys(rng, n)
    for combo, asc in ((["i16", "f64"], [True, True]), (["S7", "u8", "f32"], [True, False, True]), (["U5", "i64"], [False, True]),
                       (["b", "S20"], [True, True])):
        cols = [ks[c] for c in combo]
        pick = np.sort(rng.choice(n, 7, replace=False))
        row_offset = 1000
        perm = orc.LexSort32([np.arange(n, dtype=np.int64)] + cols, ascending=[True] + asc)
        order = np.empty(n, np.int64)
        order[perm] = np.arange(n)
        spos = np.sort(order[pick])  # sorted positions of the splitter rows
        exp = np.searchsorted(spos, order, side="right").astype(np.int32)
        srt = pick[np.argsort(order[pick])]
        got = rc.LexPartition([torch.from_numpy(c).cuda() if c.dtype.kind not in "SU" else rc._to_dev(c) for c in cols],
                              [c[srt] for c in cols], srt.astype(np.int64) + row_offset, row_offset, asc).cpu().numpy()
        np.testing.assert_array_equal(got, exp, err_msg=str(combo))


# ---- cutoffs= on the sort / pack entry points (unpinned by the reference's tests; derived semantics in the oracle) ----
def _cut_cases(n, rng):
    yield np.array([n], dtype=np.int64)
    if n >= 4:
        yield np.array([n // 3, n // 3, (2 * n) // 3, n], dtype=np.int64)  # with an empty partition
        yield np.sort(np.r_[rng.integers(0, n, 40), n]).astype(np.int64)
    if n <= 5000:
        yield np.arange(1, n + 1, dtype=np.int64)                           # one row per partition


@pytest.mark.parametrize("n", [1, 37, 4099, 300_007])
def test_lexsort_and_group_from_lexsort_with_cutoffs(rc, n):
    rng = np.random.default_rng(n)
    a = rng.integers(-3, 4, n).astype(np.int16)
    f = np.round(rng.random(n) * 5)
    f[rng.random(n) < 0.1] = np.nan
    s = rng.choice(np.array([b"", b"x", b"xy", b"zzz"], dtype="S3"), n)
    for cut in _cut_cases(n, rng):
        for keys, asc in (([a], True), ([s, f, a], True), ([f, a], [False, True])):
            got = rc.LexSort32(keys, cutoffs=cut, ascending=asc)
            exp = orc.LexSort32(keys, cutoffs=cut, ascending=asc)
            assert got.dtype == np.int32
            np.testing.assert_array_equal(got, exp)
        vals = [a, f, s]
        ig = orc.LexSort32(vals[::-1], cutoffs=cut)
        for base in (1, 0):
            gik, gifk, gcnt, gucut = rc.GroupFromLexSort(ig, vals, cutoffs=cut, base_index=base)
            eik, eifk, ecnt, eucut = orc.GroupFromLexSort(ig, vals, cutoffs=cut, base_index=base)
            np.testing.assert_array_equal(gik, eik)
            np.testing.assert_array_equal(gifk, eifk)
            np.testing.assert_array_equal(gcnt, ecnt)
            np.testing.assert_array_equal(gucut, eucut)
        # the same partitions through the hash route find the same number of groups per partition
        _, (_, hucut), hu = orc.MultiKeyGroupBy32(vals, cutoffs=cut)
        np.testing.assert_array_equal(gucut, hucut)
    with pytest.raises(ValueError):
        rc.LexSort32([a], cutoffs=np.array([n + 1], dtype=np.int64))
    if n > 1:
        with pytest.raises(ValueError):
            rc.LexSort32([a], cutoffs=np.array([n], dtype=np.int64), index=np.arange(n // 2, dtype=np.int32))


@pytest.mark.parametrize("n", [1, 50, 5001, 400_003])
@pytest.mark.parametrize("kdt", [np.int8, np.int32, np.int64])
def test_groupby_pack32_with_cutoffs(rc, n, kdt):
    rng = np.random.default_rng(n + 1)
    key = rng.integers(0, 60, n)
    for cut in _cut_cases(n, rng):
        ik, _, u = orc.MultiKeyGroupBy32([key], 0, rng.random(n) < 0.9, 2, cutoffs=cut)  # per-partition numbering, some 0s
        ik = ik.astype(kdt)
        gig, gifg, gcnt = rc.GroupByPack32(ik, None, u + 1, cutoffs=cut)
        eig, eifg, ecnt = orc.GroupByPack32(ik, None, u + 1, cutoffs=cut)
        np.testing.assert_array_equal(gig, eig)
        np.testing.assert_array_equal(gifg, eifg)
        np.testing.assert_array_equal(gcnt, ecnt)
    # without cutoffs it stays BinCount(pack=True)
    ik, _, u = orc.MultiKeyGroupBy32([key])
    gig, gifg, gcnt = rc.GroupByPack32(ik.astype(kdt), None, u + 1)
    ecnt, eig, eifg = orc.BinCount(ik, u + 1, pack=True)
    np.testing.assert_array_equal(gig, eig)
    np.testing.assert_array_equal(gifg, eifg)
    np.testing.assert_array_equal(gcnt, ecnt)


def test_riptable_glue_with_cutoffs(rc):
    """Grouping.pack_by_group's groupbypack(iKey, None, U + 1, cutoffs=pcutoffs) (riptable/rt_grouping.py:2407) and
    groupbylex(cutoffs=) as riptable's Python glue drives them (riptable/rt_numpy.py:1945-1972, 2157-2251): the rc
    module and the oracle must agree on every entry of the dicts."""
    import rt_glue

    rng = np.random.default_rng(8)
    n = 50_000
    a = rng.integers(0, 40, n).astype(np.int32)
    b = rng.choice(np.array([b"u", b"vv", b"w"], dtype="S2"), n)
    cut = np.array([10_000, 10_000, 32_123, n], dtype=np.int64)
    hk, _, hu = orc.MultiKeyGroupBy32([a, b], cutoffs=cut)
    got, exp = rt_glue.groupbypack(rc, hk, None, hu + 1, cutoffs=cut), rt_glue.groupbypack(orc, hk, None, hu + 1, cutoffs=cut)
    for k in exp:
        np.testing.assert_array_equal(got[k], exp[k], err_msg=f"groupbypack {k}")
    for fn, kw in ((rt_glue.groupbyhash, {"cutoffs": cut}), (rt_glue.groupbylex, {"cutoffs": cut})):
        got, exp = fn(rc, [a, b], **kw), fn(orc, [a, b], **kw)
        assert set(got) == set(exp)
        for k in exp:
            if exp[k] is None:
                assert got[k] is None
            elif isinstance(exp[k], tuple):
                for x, y in zip(got[k], exp[k]):
                    np.testing.assert_array_equal(x, y, err_msg=f"{fn.__name__} {k}")
            else:
                np.testing.assert_array_equal(got[k], exp[k], err_msg=f"{fn.__name__} {k}")
    assert "nUniqueCutoffs" in rt_glue.groupbylex(rc, [a, b], cutoffs=cut)


def test_lexsort64_small_and_index(rc):
    rng = np.random.default_rng(64)
    n = 200_003
    a = rng.integers(-100, 100, n).astype(np.int32)
    f = np.round(rng.random(n) * 50)
    f[rng.random(n) < 0.05] = np.nan
    got = rc.LexSort64([f, a])
    assert got.dtype == np.int64
    np.testing.assert_array_equal(got, np.lexsort([f, a]))
    idx = np.flatnonzero(rng.random(n) < 0.4).astype(np.int64)
    np.testing.assert_array_equal(rc.LexSort64([f, a], index=idx, ascending=[True, False]),
                                  orc.LexSort32([f, a], index=idx, ascending=[True, False]).astype(np.int64))
    cut = np.array([1000, 150_000, n], dtype=np.int64)
    np.testing.assert_array_equal(rc.LexSort64([f, a], cutoffs=cut), orc.LexSort32([f, a], cutoffs=cut).astype(np.int64))


def test_lexsort64_above_two_billion_rows(rc):
    """rc.LexSort64 where riptable calls it (rt_numpy.py:2669-2670): more rows than an int32 result can index.  2.2e9
    int8 keys; the int64 permutation is checked in slices: a permutation, sorted, ties in row order."""
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 100 * 2**30:
        pytest.skip("needs about 80 GB of device memory")
    n = 2_200_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    key = torch.empty(n, dtype=torch.int8, device="cuda")
    step = 1 << 28
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        key[lo:hi] = torch.randint(-100, 100, (hi - lo,), device="cuda", generator=g, dtype=torch.int8)
    perm = rc.LexSort64([key])
    assert perm.dtype == torch.int64 and perm.numel() == n
    assert int(perm.max()) == n - 1 and int(perm.min()) == 0
    seen = torch.zeros(n, dtype=torch.bool, device="cuda")
    for lo in range(0, n, step):
        seen[perm[lo:min(n, lo + step)]] = True
    assert bool(seen.all())
    del seen
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        p = perm[max(lo - 1, 0):hi]
        k = key[p]
        assert bool((k[1:] >= k[:-1]).all())
        assert bool(((k[1:] != k[:-1]) | (p[1:] > p[:-1])).all())
        del p, k
