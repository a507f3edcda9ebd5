"""Pins the CPU oracle against the reference's own golden vectors.

Each test names the reference file:line whose expected values are transcribed here.  These are the
values riptable's docstrings / tests state for riptide_cpp, so an oracle that passes them is
"parity pinned" for the behaviour they cover.
"""
import numpy as np
import pytest

from oracle import riptide_oracle as orc
import rt_glue


@pytest.fixture(scope="module")
def c2m():
    np.random.seed(12345)
    c = np.random.randint(0, 8000, 2_000_000)
    d = np.random.randint(0, 8000, 2_000_000)
    return c, d


def test_groupbyhash_docstring_plain(c2m):
    # rt_numpy.py:2013-2021
    c, _ = c2m
    r = orc.groupbyhash(c)
    assert r["unique_count"] == 8000
    assert r["iKey"][:3].tolist() == [1, 2, 3]
    assert r["iKey"][-3:].tolist() == [6061, 7889, 3002]
    assert r["iFirstKey"][:3].tolist() == [0, 1, 2]
    assert r["iFirstKey"][-3:].tolist() == [67072, 67697, 68250]
    assert r["iKey"].dtype == np.int32 and r["iFirstKey"].dtype == np.int32
    # rt_numpy.py:2037-2039
    bc = orc.BinCount(r["iKey"], r["unique_count"] + 1)
    assert bc[:3].tolist() == [0, 251, 262] and bc[-3:].tolist() == [239, 217, 246]


def test_groupbyhash_docstring_filtered(c2m):
    # rt_numpy.py:2044-2051
    c, _ = c2m
    f = (c % 3).astype(bool)
    r = orc.groupbyhash(c, filter=f)
    assert r["unique_count"] == 5333
    assert r["iKey"][:3].tolist() == [0, 1, 2]
    assert r["iKey"][-3:].tolist() == [0, 5250, 1973]
    assert r["iFirstKey"][:3].tolist() == [1, 2, 3]
    assert r["iFirstKey"][-3:].tolist() == [54422, 58655, 68250]


def test_groupbyhash_docstring_multikey(c2m):
    # rt_numpy.py:2055-2062
    c, d = c2m
    r = orc.groupbyhash([c, d])
    assert r["unique_count"] == 1968856
    assert r["iKey"][:3].tolist() == [1, 2, 3]
    assert r["iKey"][-3:].tolist() == [1968854, 1968855, 1968856]
    assert r["iFirstKey"][:3].tolist() == [0, 1, 2]
    assert r["iFirstKey"][-3:].tolist() == [1999997, 1999998, 1999999]


def test_groupbypack_docstring():
    # rt_numpy.py:1926-1933 and :2027-2033
    np.random.seed(12345)
    c = np.random.randint(0, 8, 10_000)
    x = orc.groupbyhash(c)
    assert x["iFirstKey"].tolist() == [0, 1, 3, 4, 6, 14, 18, 20]
    assert x["iKey"][:3].tolist() == [1, 2, 2] and x["iKey"][-3:].tolist() == [4, 6, 1]
    ncount = orc.BinCount(x["iKey"], x["unique_count"] + 1)
    p = orc.groupbypack(x["iKey"], ncount)
    assert p["iGroup"][:3].tolist() == [0, 9, 21] and p["iGroup"][-3:].tolist() == [9988, 9991, 9992]
    assert p["iFirstGroup"].tolist() == [0, 0, 1213, 2465, 3761, 4987, 6239, 7522, 8797]
    assert p["nCountGroup"].tolist() == [0, 1213, 1252, 1296, 1226, 1252, 1283, 1275, 1203]
    assert p["nCountGroup"].sum() == 10000
    p2 = orc.groupbypack(x["iKey"], None, x["unique_count"] + 1)
    for k in p:
        assert np.array_equal(p[k], p2[k])


GBLEX_IKEY = [0, 1, 9, 0, 23, 31, 0, 45, 53, 0, 2, 3, 0, 4, 5, 0, 6, 7, 0, 8, 10, 0, 11, 12, 0, 13, 14, 0, 15,
              16, 0, 17, 18, 0, 19, 20, 0, 21, 22, 0, 24, 25, 0, 26, 27, 0, 28, 29, 0, 30, 32, 0, 33, 34, 0, 35,
              36, 0, 37, 38, 0, 39, 40, 0, 41, 42, 0, 43, 44, 0, 46, 47, 0, 48, 49, 0, 50, 51, 0, 52, 54, 0, 55,
              56, 0, 57, 58, 0, 59, 60, 0, 61, 62, 0, 63, 64, 0, 65, 66, 0]
GBLEX_IFIRST = [1, 10, 11, 13, 14, 16, 17, 19, 2, 20, 22, 23, 25, 26, 28, 29, 31, 32, 34, 35, 37, 38, 4, 40, 41,
                43, 44, 46, 47, 49, 5, 50, 52, 53, 55, 56, 58, 59, 61, 62, 64, 65, 67, 68, 7, 70, 71, 73, 74, 76,
                77, 79, 8, 80, 82, 83, 85, 86, 88, 89, 91, 92, 94, 95, 97, 98]


def test_groupbylex_docstring():
    # rt_numpy.py:2126-2156
    a = np.arange(100).astype("S")
    f = (np.arange(100) % 3).astype(bool)
    r = orc.groupbylex([a], filter=f)
    assert r["iKey"].tolist() == GBLEX_IKEY
    assert r["iFirstKey"].tolist() == GBLEX_IFIRST
    assert r["unique_count"] == 66
    assert r["iGroup"].tolist() == GBLEX_IFIRST
    assert r["iFirstGroup"].tolist() == [66] + list(range(66))
    assert r["nCountGroup"].tolist() == [34] + [1] * 66


def test_combine_accum1_filter_docstring():
    # rt_numpy.py:1533-1540: Cat of (arange(20)%10).astype('S'), ordered -> ikey = value+1
    key = ((np.arange(20) % 10) + 1).astype(np.int8)
    ik, ifk, u = orc.CombineAccum1Filter(key, 10, (np.arange(20) % 2).astype(bool))
    assert ik.tolist() == [0, 1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 1, 0, 2, 0, 3, 0, 4, 0, 5]
    assert ik.dtype == np.int8
    assert ifk.tolist() == [1, 3, 5, 7, 9] and u == 5


def test_reductions_kat_test_groupby():
    # tests/test_groupby.py:13-104: key = bool [T,F,T,F,T]; ds.gb sorts keys -> bins (False, True)
    num = [3, 4, 5, 1, 2]
    key = np.array([True, False, True, False, True])
    ikey = np.where(key, 2, 1).astype(np.int8)  # sorted display order: False=1, True=2
    for dt in [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32,
               np.float64]:
        v = np.array(num, dtype=dt)

        def run(fn):
            return orc.GroupByAll32([v], ikey, 2, [fn], [1], [3], 0)[0][1:]

        assert run(orc.GB_SUM).tolist() == [5, 10]
        assert run(orc.GB_NANSUM).tolist() == [5, 10]
        assert run(orc.GB_MIN).tolist() == [1, 2] and run(orc.GB_MIN).dtype == dt
        assert run(orc.GB_MAX).tolist() == [4, 5]
        np.testing.assert_allclose(run(orc.GB_MEAN), [2.5, 3.333333333], rtol=1e-6)
        np.testing.assert_allclose(run(orc.GB_STD), [2.121320343560, 1.527525231652], rtol=1e-6)
        np.testing.assert_allclose(run(orc.GB_VAR), [4.5, 2.333333333333], rtol=1e-6)
    # tests/test_groupby.py:191-210: var/std of a 1-item group is NaN
    ik1 = np.arange(1, 6).astype(np.int8)
    for fn in (orc.GB_VAR, orc.GB_STD, orc.GB_NANVAR, orc.GB_NANSTD):
        assert np.isnan(orc.GroupByAll32([np.array(num, float)], ik1, 5, [fn], [1], [6], 0)[0][1:]).all()
    # tests/test_groupby.py:212-225 count
    assert orc.BinCount(ikey, 3)[1:].tolist() == [2, 3]


@pytest.mark.parametrize("N,N_cat", [(1, 1), (5, 3), (100, 10), (3467, 1000), (60_000, 10)])
@pytest.mark.parametrize("nan_fraction", [0.0, 0.3, 0.9999999, 1.0])
def test_min_max_replays_reference_test(N, N_cat, nan_fraction):
    # tests/test_groupbyops.py:389-456: min / max propagate NaN (and, for integers, the sentinel), nanmin / nanmax skip them;
    # compared per category with np.min / np.max / np.nanmin / np.nanmax on a float copy with NaN at the sentinels
    import warnings
    rng = np.random.default_rng((N * N_cat * int(nan_fraction * 1000)) % 10_000)
    cat = rng.integers(0, N_cat, size=N)
    g = orc.groupbyhash(cat)
    ik, U = g["iKey"], g["unique_count"]
    data = rng.normal(rng.uniform(-10_000, 10_000), rng.uniform(0.0, 10_000), size=N)
    data[rng.random(N) < nan_fraction] = np.nan
    idata = rng.integers(-10_000, 10_000, size=N)
    inan = rng.random(N) < nan_fraction
    idata[inan] = np.iinfo(idata.dtype).min
    iref = idata.astype(np.float64)
    iref[inan] = np.nan
    for col, ref in ((data, data), (idata, iref)):
        res = {fn: orc.GroupByAll32([col], ik, U, [fn], [1], [U + 1], 0)[0] for fn in (orc.GB_MIN, orc.GB_MAX, orc.GB_NANMIN, orc.GB_NANMAX)}
        for fn in res:
            assert res[fn].dtype == col.dtype
        for b in range(1, U + 1):
            sl = ref[ik == b]
            with np.errstate(all="ignore"), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                exp = {orc.GB_MIN: np.min(sl), orc.GB_MAX: np.max(sl), orc.GB_NANMIN: np.nanmin(sl), orc.GB_NANMAX: np.nanmax(sl)}
            for fn, e in exp.items():
                got = float(res[fn][b])
                if col.dtype.kind == "i" and res[fn][b] == np.iinfo(col.dtype).min:
                    got = np.nan  # the integer sentinel reads as NaN (FastArray.astype(float))
                assert (np.isnan(e) and np.isnan(got)) or got == e, (fn, b, got, e)


def test_mean_int64_equals_mean_float64_and_empty_bins():
    # tests/test_groupby.py:393-406: mean of arange(300000) as int64 equals the mean of the same values as float64
    k = (np.arange(300_000) % 3 + 1).astype(np.int8)
    mi = orc.GroupByAll32([np.arange(300_000)], k, 3, [orc.GB_MEAN], [1], [4], 0)[0]
    mf = orc.GroupByAll32([np.arange(300_000.0)], k, 3, [orc.GB_MEAN], [1], [4], 0)[0]
    assert mi.dtype == np.float64 and (mi[1:] == mf[1:]).all()
    assert orc.GroupByAll32([(np.arange(300_000) % 3)], k, 3, [orc.GB_MEAN], [1], [4], 0)[0][2] == 1.0
    # tests/test_categorical_groupby.py:776-808: an empty category gives 0 for sum / nansum and NaN for everything else
    kk = np.array([1, 2, 4, 5, 1, 2], np.int8)  # bin 3 never occurs
    v = np.arange(6.0)
    for fn, exp in ((orc.GB_SUM, 0.0), (orc.GB_NANSUM, 0.0), (orc.GB_MEAN, np.nan), (orc.GB_MIN, np.nan), (orc.GB_MAX, np.nan),
                    (orc.GB_VAR, np.nan), (orc.GB_STD, np.nan), (orc.GB_NANMEAN, np.nan), (orc.GB_NANMIN, np.nan),
                    (orc.GB_NANMAX, np.nan), (orc.GB_NANVAR, np.nan), (orc.GB_NANSTD, np.nan)):
        r = orc.GroupByAll32([v], kk, 5, [fn], [1], [6], 0)[0][3]
        assert (np.isnan(exp) and np.isnan(r)) or r == exp, fn


def test_unique_semantics_of_the_grouping():
    # tests/test_unique.py:14-48: grouping + iFirstKey reproduce np.unique (items, first index, inverse, counts) for every
    # numeric dtype, viewing the same random bytes (so denormals, negative zeros and NaN bit patterns are all there for
    # floats -- keys compare bytewise: distinct bit patterns are distinct keys)
    rng = np.random.default_rng(9)
    raw = rng.integers(0, 256, 8 * 4000, dtype=np.uint8)
    raw[: 8 * 2000] = raw[8 * 2000:]  # every value appears at least twice
    for dt in "bBhHiIlL":
        nums = raw.view(np.dtype(dt))
        ik, ifk, U = orc.MultiKeyGroupBy32([nums], 0, None, 2)
        np_un, np_idx, np_inv, np_cnt = np.unique(nums, return_index=True, return_inverse=True, return_counts=True)
        order = np.argsort(nums[ifk], kind="stable")  # hash bins are in first-appearance order; np.unique sorts
        assert U == len(np_un)
        np.testing.assert_array_equal(nums[ifk][order], np_un)
        np.testing.assert_array_equal(ifk[order], np_idx)
        rank = np.empty(U, np.int64)
        rank[order] = np.arange(U)
        np.testing.assert_array_equal(rank[ik - 1], np_inv)
        np.testing.assert_array_equal(orc.BinCount(ik, U + 1)[1:][order], np_cnt)
    for dt in "fd":
        bits = raw.view(np.dtype(dt).str.replace("f", "u"))
        ik, _, U = orc.MultiKeyGroupBy32([raw.view(np.dtype(dt))], 0, None, 2)
        ikb, _, Ub = orc.MultiKeyGroupBy32([bits], 0, None, 2)
        assert U == Ub and np.array_equal(ik, ikb)


def test_reductions_against_pandas():
    # tests/test_groupby_autotest_aggregated_functions.py:26-29, 566-584: sum / mean / min / max / var / std (ddof = 1)
    # / first / last against pandas' groupby, to 7 places there; exact or 1e-12 here
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(77)
    n = 20_000
    key = rng.integers(0, 60, n)
    v = rng.normal(5, 100, n)
    vi = rng.integers(-1000, 1000, n)
    g = orc.groupbyhash(key)
    ik, U = g["iKey"], g["unique_count"]
    uniq = key[g["iFirstKey"]]
    df = pd.DataFrame({"k": key, "v": v, "vi": vi})
    gb = df.groupby("k")

    def ours(col, fn):
        return pd.Series(orc.GroupByAll32([col], ik, U, [fn], [1], [U + 1], 0)[0][1:], index=uniq).sort_index().to_numpy()

    for fn, name in ((orc.GB_SUM, "sum"), (orc.GB_MEAN, "mean"), (orc.GB_MIN, "min"), (orc.GB_MAX, "max"), (orc.GB_VAR, "var"),
                     (orc.GB_STD, "std")):
        np.testing.assert_allclose(ours(v, fn), getattr(gb["v"], name)().to_numpy(), rtol=1e-11, err_msg=name)
    for fn, name in ((orc.GB_SUM, "sum"), (orc.GB_MIN, "min"), (orc.GB_MAX, "max")):
        np.testing.assert_array_equal(ours(vi, fn), getattr(gb["vi"], name)().to_numpy(), err_msg=name)
    np.testing.assert_allclose(ours(vi, orc.GB_MEAN), gb["vi"].mean().to_numpy(), rtol=1e-13)
    cnt, ig, ifg = orc.BinCount(ik, U + 1, pack=True)
    first = orc.GroupByAllPack32([v], ik, ig, ifg, cnt, U, [orc.GB_FIRST], [1], [U + 1], 0)[0][1:]
    last = orc.GroupByAllPack32([v], ik, ig, ifg, cnt, U, [orc.GB_LAST], [1], [U + 1], 0)[0][1:]
    np.testing.assert_array_equal(pd.Series(first, index=uniq).sort_index().to_numpy(), gb["v"].first().to_numpy())
    np.testing.assert_array_equal(pd.Series(last, index=uniq).sort_index().to_numpy(), gb["v"].last().to_numpy())
    np.testing.assert_array_equal(pd.Series(orc.BinCount(ik, U + 1)[1:], index=uniq).sort_index().to_numpy(), gb.size().to_numpy())


def test_accum2_simple_cats_kat():
    # tests/test_accum2.py:87-127: Accum2 of two 5-category Categoricals over data = 10, 20 .. 50 puts data[i] on the
    # diagonal; with an operation filter the filtered rows fall into cell 0
    data = np.arange(1, 6) * 10
    k1 = np.arange(1, 6, dtype=np.int8)  # Categorical(["a" .. "e"]) codes
    k2 = np.arange(1, 6, dtype=np.int8)  # Categorical(arange(5)) codes
    ik, cnt = orc.CombineAccum2Filter(k1, k2, 5, 5, None)
    nb = 6 * 6
    assert len(cnt) == nb
    cells = orc.GroupByAll32([data], ik, nb - 1, [orc.GB_SUM], [0], [nb], 0)[0].reshape(6, 6)
    for i in range(5):
        assert cells[i + 1, i + 1] == data[i]
    assert cells.sum() == data.sum() and cnt.reshape(6, 6)[1:, 1:].trace() == 5
    f = np.array([True, False, True, True, False])
    ikf, cntf = orc.CombineAccum2Filter(k1, k2, 5, 5, f)
    cells = orc.GroupByAll32([data], ikf, nb - 1, [orc.GB_SUM], [0], [nb], 0)[0].reshape(6, 6)
    assert cells[0, 0] == 20 + 50 and cells[2, 2] == 0 and cells[1, 1] == 10 and cntf[0] == 2


def test_first_last_kat():
    # tests/test_groupby.py:137-159: first [4,3], last [1,2] for key bool sorted (False, True)
    num = np.array([3, 4, 5, 1, 2])
    ikey = np.array([2, 1, 2, 1, 2], dtype=np.int8)
    cnt, ig, ifg = orc.BinCount(ikey, 3, pack=True)
    first = orc.GroupByAllPack32([num], ikey, ig, ifg, cnt, 2, [orc.GB_FIRST], [1], [3], 0)[0]
    last = orc.GroupByAllPack32([num], ikey, ig, ifg, cnt, 2, [orc.GB_LAST], [1], [3], 0)[0]
    assert first[1:].tolist() == [4, 3] and last[1:].tolist() == [1, 2]
    s = np.array([b"3", b"4", b"5", b"1", b"2"])
    assert orc.GroupByAllPack32([s], ikey, ig, ifg, cnt, 2, [orc.GB_FIRST], [1], [3], 0)[0][1:].tolist() == [b"4", b"3"]
    # nth out of range -> invalid (tests/test_groupby.py:755-770)
    nth = orc.GroupByAllPack32([num.astype(float)], ikey, ig, ifg, cnt, 2, [orc.GB_NTH], [1], [3], 2)[0]
    assert np.isnan(nth[1]) and nth[2] == 2.0
    nthn = orc.GroupByAllPack32([num], ikey, ig, ifg, cnt, 2, [orc.GB_NTH], [1], [3], -3)[0]
    assert nthn[1] == np.iinfo(num.dtype).min and nthn[2] == 3


def test_nth_offsets_and_filter_kat():
    # tests/test_categorical_groupby.py:1107-1162: nth(n) for n in -3..2 on groups of 1 and 2 rows, numbers and strings;
    # out of range -> NaN / empty string; nth with an operation filter (filtered rows fall into bin 0 and are found there
    # with showfilter)
    ik = np.array([1, 2, 2], np.int8)
    num = np.array([123.0, 456.0, 789.0])
    st = np.array([b"Aa", b"Bb", b"Cc"])
    cnt, ig, ifg = orc.BinCount(ik, 3, pack=True)
    offs = [0, 1]
    for n in range(-3, 3):
        rn, rs = orc.GroupByAllPack32([num, st], ik, ig, ifg, cnt, 2, [orc.GB_NTH] * 2, [1, 1], [3, 3], n)
        for grp in range(2):
            gcount = grp + 1
            gi = n if n >= 0 else gcount + n
            if 0 <= gi < gcount:
                assert rn[grp + 1] == num[offs[grp] + gi] and rs[grp + 1] == st[offs[grp] + gi]
            else:
                assert np.isnan(rn[grp + 1]) and len(rs[grp + 1]) == 0
    f = np.array([True, False, True])
    ikf = orc.CombineFilter(ik, f)
    cnt, ig, ifg = orc.BinCount(ikf, 3, pack=True)
    rn, rs = orc.GroupByAllPack32([num, st], ikf, ig, ifg, cnt, 2, [orc.GB_NTH] * 2, [0, 0], [3, 3], 0)
    assert rn.tolist() == [456.0, 123.0, 789.0] and rs.tolist() == [b"Bb", b"Aa", b"Cc"]  # slot 0 = the filtered rows


def test_cumsum_kat():
    # tests/test_groupby.py:504-602
    order_ids = np.array([1, 1, 2, 1, 2, 2, 1, 2, 2, 2, 2, 1])
    seconds = np.array([50, 70, 72, 75, 90, 88, 95, 97, 98, 115, 116, 120])
    shares = np.array([0, 200, 0, 500, 100, 400, 100, 0, 300, 150, 150, 0])
    r = orc.groupbyhash(order_ids)
    ik = r["iKey"]
    out = orc.EmaAll32([shares, seconds], ik, 2, orc.GB_CUMSUM, (0.0, None, None, None))
    assert out[0].tolist() == [0, 200, 0, 700, 100, 500, 800, 500, 800, 950, 1100, 800]
    assert out[1].tolist() == [50, 120, 72, 195, 162, 250, 290, 347, 445, 560, 676, 410]
    f = np.arange(12) % 2 == 0
    out = orc.EmaAll32([shares, seconds], ik, 2, orc.GB_CUMSUM, (0.0, None, f, None))
    assert out[0].tolist() == [0, 0, 0, 0, 100, 100, 100, 100, 400, 400, 550, 100]
    assert out[1].tolist() == [50, 50, 72, 50, 162, 162, 145, 162, 260, 260, 376, 145]
    # filter at construction: bin-0 rows -> invalid of the output dtype
    rf = orc.groupbyhash(order_ids, filter=f)
    out = orc.EmaAll32([shares, seconds], rf["iKey"], rf["unique_count"], orc.GB_CUMSUM, (0.0, None, None, None))
    inv = np.iinfo(np.int64).min
    assert out[0].dtype == np.int64
    assert out[0].tolist() == [0, inv, 0, inv, 100, inv, 100, inv, 400, inv, 550, inv]
    assert out[1].tolist() == [50, inv, 72, inv, 162, inv, 145, inv, 260, inv, 376, inv]


def _as_float_with_nan(a):
    """FastArray.astype(float): integer sentinels become NaN (what the reference test compares in)."""
    out = a.astype(np.float64)
    if a.dtype.kind in "iu":
        out[a == orc.invalid_for(a.dtype)] = np.nan
    return out


@pytest.mark.parametrize("N,N_cat", [(1, 1), (5, 10), (97, 10), (2643, 45)])
@pytest.mark.parametrize("nan_fraction", [0.0, 0.5, 1.0])
@pytest.mark.parametrize("skipna", [True, False])
def test_cum_min_max_replays_reference_test(N, N_cat, nan_fraction, skipna):
    # tests/test_groupbyops.py:458-538, with numpy in place of Dataset / Categorical
    rng = np.random.default_rng((N * N_cat * int(nan_fraction * 1000)) % 10_000)
    categ = rng.integers(0, N_cat, size=N)
    g = orc.groupbyhash(categ)
    ik, U = g["iKey"], g["unique_count"]
    cols = []
    for dt in [np.float32, np.float64]:
        d = rng.normal(rng.uniform(-10_000, 10_000), rng.uniform(0.0, 10_000), size=N).astype(dt)
        if nan_fraction > 0:
            d[rng.random(N) < nan_fraction / 3] = np.nan
            d[rng.random(N) < nan_fraction / 3] = np.inf
            d[rng.random(N) < nan_fraction / 3] = -np.inf
        cols.append(d)
    for dt in [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64]:
        d = rng.integers(-10_000, 10_000, size=N).astype(dt)
        d[rng.random(N) < nan_fraction] = orc.invalid_for(dt)
        cols.append(d)
    cols.append(rng.choice([True, False], size=N))
    fmin, fmax = (orc.GB_CUMNANMIN, orc.GB_CUMNANMAX) if skipna else (orc.GB_CUMMIN, orc.GB_CUMMAX)
    cmin = orc.EmaAll32(cols, ik, U, fmin, (0.0, None, None, None))
    cmax = orc.EmaAll32(cols, ik, U, fmax, (0.0, None, None, None))
    for c, lo, hi in zip(cols, cmin, cmax):
        assert lo.dtype == c.dtype and hi.dtype == c.dtype
        for b in range(1, U + 1):
            sl = _as_float_with_nan(c[ik == b])
            with np.errstate(invalid="ignore"):
                if skipna:
                    elo, ehi = np.fmin.accumulate(sl), np.fmax.accumulate(sl)
                else:
                    elo, ehi = np.minimum.accumulate(sl), np.maximum.accumulate(sl)
            np.testing.assert_array_equal(_as_float_with_nan(lo[ik == b]), elo)
            np.testing.assert_array_equal(_as_float_with_nan(hi[ik == b]), ehi)


def test_cum_ops_against_pandas():
    # tests/test_categorical_functions.py / test_groupby_functions.py compare the cumulative functions with pandas' groupby
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(12)
    n = 5000
    key = rng.integers(0, 17, n)
    g = orc.groupbyhash(key)
    ik, U = g["iKey"], g["unique_count"]
    v = rng.normal(0, 3, n)
    vp = np.exp(rng.normal(0, 0.05, n))
    df = pd.DataFrame({"k": key, "v": v, "vp": vp})
    none = (0.0, None, None, None)
    np.testing.assert_allclose(orc.EmaAll32([v], ik, U, orc.GB_CUMSUM, none)[0], df.groupby("k")["v"].cumsum().to_numpy(), rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(orc.EmaAll32([vp], ik, U, orc.GB_CUMPROD, none)[0], df.groupby("k")["vp"].cumprod().to_numpy(), rtol=1e-12)
    np.testing.assert_array_equal(orc.EmaAll32([v], ik, U, orc.GB_CUMNANMIN, none)[0], df.groupby("k")["v"].cummin().to_numpy())
    np.testing.assert_array_equal(orc.EmaAll32([v], ik, U, orc.GB_CUMNANMAX, none)[0], df.groupby("k")["v"].cummax().to_numpy())
    vi = rng.integers(-9, 9, n)
    np.testing.assert_array_equal(orc.EmaAll32([vi], ik, U, orc.GB_CUMSUM, none)[0], pd.Series(vi).groupby(key).cumsum().to_numpy())


def test_cumprod_small():
    ik = np.array([1, 2, 1, 0, 2, 1], dtype=np.int8)
    v = np.array([2, 3, 4, 9, 5, 6], dtype=np.int32)
    out = orc.EmaAll32([v, v.astype(np.float32)], ik, 2, orc.GB_CUMPROD, (0.0, None, None, None))
    assert out[0].dtype == np.int64 and out[0].tolist() == [2, 3, 8, np.iinfo(np.int64).min, 15, 48]
    assert out[1].dtype == np.float32 and out[1][[0, 1, 2, 4, 5]].tolist() == [2.0, 3.0, 8.0, 15.0, 48.0] and np.isnan(out[1][3])
    f = np.array([True, True, False, True, True, True])
    assert orc.EmaAll32([v], ik, 2, orc.GB_CUMPROD, (0.0, None, f, None))[0].tolist() == [2, 3, 2, np.iinfo(np.int64).min, 15, 12]


def test_rolling_ops_pinned_to_pandas_and_kat():
    # tests/test_groupby.py:652-708 compare shift / diff with pandas; tests/test_categorical_groupby.py:749-763 cumcount;
    # tests/test_fastarray.py:1297-1305 the partial-window nansum KAT
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(0)
    n = 300
    key = rng.choice(np.array(["a", "b", "c"]), n)
    g = orc.groupbyhash(key)
    ik, U = g["iKey"], g["unique_count"]
    cnt, ig, ifg = orc.BinCount(ik, U + 1, pack=True)
    v = rng.random(n)
    v[rng.random(n) < 0.1] = np.nan
    df = pd.DataFrame({"k": key, "v": v})
    call = lambda vals, fn, p: orc.GroupByAllPack32([vals], ik, ig, ifg, cnt, U, [fn], [0], [U + 1], p)[0]
    for per in (-2, 3, 1):
        np.testing.assert_array_equal(call(v, orc.GB_ROLLING_SHIFT, per), df.groupby("k")["v"].shift(per).to_numpy())
        np.testing.assert_allclose(call(v, orc.GB_ROLLING_DIFF, per), df.groupby("k")["v"].diff(per).to_numpy(), rtol=0, atol=0)
    for w in (1, 3, 7):
        roll = lambda mp: df.groupby("k")["v"].rolling(w, min_periods=mp)
        flat = lambda r: r.reset_index(level=0, drop=True).sort_index().to_numpy()
        np.testing.assert_allclose(call(v, orc.GB_ROLLING_NANSUM, w), np.nan_to_num(flat(roll(0).sum())), rtol=1e-12)
        np.testing.assert_allclose(call(v, orc.GB_ROLLING_NANMEAN, w), flat(roll(1).mean()), rtol=1e-12)
    np.testing.assert_array_equal(call(ik, orc.GB_ROLLING_COUNT, 1), df.groupby("k").cumcount().to_numpy())
    np.testing.assert_array_equal(call(ik, orc.GB_ROLLING_COUNT, -1), df.groupby("k").cumcount(ascending=False).to_numpy())
    # cumcount with filtered rows: invalid there (test_categorical_groupby.py:760-763)
    f = np.arange(n) % 2 == 1
    ikf = orc.CombineFilter(ik, f)
    cf, igf, iff = orc.BinCount(ikf, U + 1, pack=True)
    cc = orc.GroupByAllPack32([ikf], ikf, igf, iff, cf, U, [orc.GB_ROLLING_COUNT], [0], [U + 1], 1)[0]
    assert (cc[~f] == np.iinfo(np.int32).min).all() and (cc[f] >= 0).all()
    x = np.array([1, 2, 3, np.nan, np.nan, np.nan])
    one = np.ones(6, np.int8)
    c1, g1, f1 = orc.BinCount(one, 2, pack=True)
    assert orc.GroupByAllPack32([x], one, g1, f1, c1, 1, [orc.GB_ROLLING_NANSUM], [0], [2], 3)[0].tolist() == [1, 3, 6, 5, 3, 0]
    xi = np.array([1, 2, 3, 0, 0, 0], np.int64)
    xi[3:] = np.iinfo(np.int64).min
    r = orc.GroupByAllPack32([xi], one, g1, f1, c1, 1, [orc.GB_ROLLING_NANSUM], [0], [2], 3)[0]
    assert r.dtype == np.int64 and r.tolist() == [1, 3, 6, 5, 3, 0]


@pytest.mark.parametrize("N,N_cat,q,nan_fraction,dtype", [
    (1000, 7, 0.5, 0.1, np.float64), (5000, 45, 0.29, 0.0, np.int32), (97, 1, 0.9, 0.3, np.float64),
    (2643, 10, 0.0, 0.2, np.int32), (500, 3, 1.0, 0.5, np.float64), (10_000, 100, 0.123, 0.05, np.float64)])
@pytest.mark.parametrize("with_infs", [False, True])
def test_quantile_replays_reference_test(N, N_cat, q, nan_fraction, dtype, with_infs):
    # tests/test_groupbyops.py:540-700: quantile / nanquantile against np.quantile / np.nanquantile ("midpoint"),
    # infinities through (lower + higher) / 2 (rt_numpy.py:4202-4245); integer columns carry sentinels for NaN
    import warnings
    rng = np.random.default_rng((N * N_cat * int(q * 1000) * int(nan_fraction * 1000)) % 10_000)
    cat = rng.integers(0, N_cat, N)
    g = orc.groupbyhash(cat)
    ik, U = g["iKey"], g["unique_count"]
    cnt, ig, ifg = orc.BinCount(ik, U + 1, pack=True)
    data = rng.normal(rng.uniform(-10_000, 10_000), rng.uniform(0.0, 10_000), size=N).astype(dtype)
    ref = data.astype(np.float64)
    if with_infs and np.dtype(dtype).kind == "f":
        for val in (np.inf, -np.inf):
            m = rng.random(N) < 0.1
            data[m] = val
            ref[m] = val
    nan_idx = rng.random(N) < nan_fraction
    ref[nan_idx] = np.nan
    data[nan_idx] = orc.invalid_for(dtype)
    for is_nan, npf in ((False, np.quantile), (True, np.nanquantile)):
        got = orc.GroupByAllPack32([data], ik, ig, ifg, cnt, U, [orc.GB_QUANTILE_MULT], [1], [U + 1], orc.quantile_func_param(q, is_nan))[0]
        assert got.dtype == np.float64 and np.isnan(got[0])
        for b in range(1, U + 1):
            sl = ref[ik == b]
            with np.errstate(all="ignore"), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                exp = (npf(sl, q, method="lower") + npf(sl, q, method="higher")) / 2
            np.testing.assert_allclose(got[b], exp, rtol=1e-12, equal_nan=True)
    # GB_MEDIAN "auto handles nan" (rt_enum.py:506); tests/test_groupby.py:83 median KAT
    med = orc.GroupByAllPack32([data], ik, ig, ifg, cnt, U, [orc.GB_MEDIAN], [1], [U + 1], 0)[0]
    nanmed = orc.GroupByAllPack32([data], ik, ig, ifg, cnt, U, [orc.GB_QUANTILE_MULT], [1], [U + 1], orc.quantile_func_param(0.5, True))[0]
    np.testing.assert_array_equal(med, nanmed)
    # tests/test_groupby.py:13-104: key bool [T,F,T,F,T] (bins False=1, True=2), values [3,4,5,1,2] -> medians [2.5, 3]
    kb = np.array([2, 1, 2, 1, 2], np.int8)
    c2, g2, f2 = orc.BinCount(kb, 3, pack=True)
    for vdt in (np.int8, np.uint16, np.int64, np.float32, np.float64):
        vals = np.array([3, 4, 5, 1, 2], vdt)
        assert orc.GroupByAllPack32([vals], kb, g2, f2, c2, 2, [orc.GB_MEDIAN], [1], [3], 0)[0][1:].tolist() == [2.5, 3.0]


def test_ema_docstring_kats():
    # rt_groupbyops.py:3378-3386 (ema_decay; the table shows its inputs rounded to 2 decimals) and :3424-3446 /
    # :3486-3508 (ema_normal / ema_weighted on rt.arange(10) grouped by arange(10) % 3, printed to 2 decimals)
    ik = np.ones(3, np.int8)
    r = orc.EmaAll32([np.array([-3.11, 210.54, 49.97])], ik, 1, orc.GB_EMADECAY,
                     (np.log(2) / (1e3 * 100), np.array([25.65, 38.37, 41.66]), None, None))[0]
    np.testing.assert_allclose(r, [-3.11271882, 207.42784495, 257.39155897], rtol=1e-3)
    test = np.arange(10)
    g = (np.arange(10) % 3 + 1).astype(np.int8)
    normal = orc.EmaAll32([test], g, 3, orc.GB_EMANORMAL, (1.0, np.arange(10), None, None))[0]
    weighted = orc.EmaAll32([test], g, 3, orc.GB_EMAWEIGHTED, (0.5, None, None, None))[0]
    assert np.round(normal, 2).tolist() == [0.00, 1.00, 2.00, 2.85, 3.85, 4.85, 5.84, 6.84, 7.84, 8.84]
    assert np.round(weighted, 2).tolist() == [0.00, 1.00, 2.00, 1.50, 2.50, 3.50, 3.75, 4.75, 5.75, 6.38]
    # this restatement's choices for what the reference leaves open: excluded rows show the running value
    f = np.array([True, True, True, False, True, True, True, True, True, True])
    w = orc.EmaAll32([test], g, 3, orc.GB_EMAWEIGHTED, (0.5, None, f, None))[0]
    assert w[3] == 0.0 and w[6] == 3.0  # row 3 excluded: group 1 still at 0; row 6: 6 * .5 + 0 * .5


def test_ismember_categorical_ordered_remap():
    # rt_grouping.py:955-963: an ordered Categorical turns first-appearance codes into sorted-order codes with
    # ismember(tempikey - 1, sortidx, base_index=1); a filtered row (code 0 -> -1) stays 0
    keys = np.array(["pear", "apple", "fig", "apple", "pear", "kiwi"])
    ik, ifk, U = orc.MultiKeyGroupBy32([keys], 0, np.array([True, True, True, True, True, False]), 2)
    uniq = keys[ifk]
    sortidx = np.argsort(uniq, kind="stable")
    found, new = orc.IsMemberCategorical(ik.astype(np.int32) - 1, sortidx.astype(np.int32))
    assert new.tolist() == [3, 1, 2, 1, 3, 0] and found.tolist() == [True] * 5 + [False]
    assert uniq[sortidx][new[:5] - 1].tolist() == keys[:5].tolist()


@pytest.mark.parametrize("N,N_cat,window,q,nan_fraction,dtype", [
    (500, 1, 150, 0.5, 0.1, np.float64), (2000, 40, 1, 0.3, 0.2, np.float64), (300, 35, 800, 0.75, 0.0, np.int32),
    (1500, 12, 3, 0.0, 0.3, np.float64), (1500, 12, 17, 1.0, 0.05, np.int32), (800, 5, 25, 0.123, 1.0, np.float64)])
def test_rolling_quantile_replays_reference_test(N, N_cat, window, q, nan_fraction, dtype):
    # tests/test_groupbyops.py:901-963: from the first full window on, every row equals the nan-quantile (midpoint) of the
    # last `window` rows of its category; window is capped by the category's size
    rng = np.random.default_rng((N * N_cat * int(q * 1000) * int(nan_fraction * 1000)) % 10_000)
    cat = rng.integers(0, N_cat, N)
    g = orc.groupbyhash(cat)
    ik, U = g["iKey"], g["unique_count"]
    cnt, ig, ifg = orc.BinCount(ik, U + 1, pack=True)
    data = rng.normal(rng.uniform(-100, 100), rng.uniform(0.0, 100), size=N).astype(dtype)
    ref = data.astype(np.float64)
    nan_idx = rng.random(N) < nan_fraction
    ref[nan_idx] = np.nan
    data[nan_idx] = orc.invalid_for(dtype)
    got = orc.GroupByAllPack32([data], ik, ig, ifg, cnt, U, [orc.GB_ROLLING_QUANTILE], [0], [U + 1],
                               orc.rolling_quantile_func_param(q, window))[0]
    assert got.dtype == np.float64 and len(got) == N
    for b in range(1, U + 1):
        sl, res = ref[ik == b], got[ik == b]
        w = min(window, len(sl))
        for p in range(w - 1, len(sl)):
            vals = np.sort(sl[p - w + 1: p + 1][~np.isnan(sl[p - w + 1: p + 1])])
            if len(vals) == 0:
                assert np.isnan(res[p])
                continue
            idx = (len(vals) - 1) * q
            exp = (vals[int(np.floor(idx))] + vals[int(np.ceil(idx))]) / 2
            assert res[p] == exp, (b, p, res[p], exp)


def test_mode_kat():
    # tests/test_groupby.py:838-861: key [a, b, b, b], v_int [1, 1, 2, 2] -> mode per key [1, 2]
    ik = np.array([1, 2, 2, 2], np.int8)
    cnt, ig, ifg = orc.BinCount(ik, 3, pack=True)
    for dt in (np.int32, np.int64, np.float64):
        v = np.array([1, 1, 2, 2], dt)
        r = orc.GroupByAllPack32([v], ik, ig, ifg, cnt, 2, [orc.GB_MODE], [1], [3], 0)[0]
        assert r.dtype == dt and r[1:].tolist() == [1, 2]
    v = np.array([np.nan, 3.0, np.nan, np.nan])
    r = orc.GroupByAllPack32([v], ik, ig, ifg, cnt, 2, [orc.GB_MODE], [1], [3], 0)[0]
    assert np.isnan(r[1]) and r[2] == 3.0  # "auto handles nan": NaN never counts


def test_cutoffs_kat():
    # tests/test_pgroupby.py:7-86
    rng = np.random.default_rng(0)
    arr, arr2 = rng.integers(0, 5, 15), rng.integers(0, 5, 15)
    arr3 = np.concatenate([arr, arr2])
    g1, g2 = orc.groupbyhash([arr]), orc.groupbyhash([arr2])
    g3 = orc.groupbyhash([arr3], cutoffs=np.array([15, 30], dtype=np.int64))
    vals, cuts = g3["iFirstKey"]
    assert np.array_equal(g1["iKey"], g3["iKey"][:15]) and np.array_equal(g2["iKey"], g3["iKey"][15:])
    assert np.array_equal(g1["iFirstKey"], vals[: cuts[0]]) and np.array_equal(g2["iFirstKey"], vals[cuts[0]:])
    assert g3["unique_count"] == g1["unique_count"] + g2["unique_count"]
    mk = np.concatenate([np.zeros(15), np.ones(15)])
    gm = orc.groupbyhash([mk, arr3])
    assert np.array_equal(gm["iFirstKey"][: cuts[0]], vals[: cuts[0]])
    assert np.array_equal(gm["iFirstKey"][cuts[0]:], vals[cuts[0]:] + 15)
    # test_empty_cutoffs
    a = np.arange(10)
    pgb = orc.groupbyhash([np.ones(10)], filter=(a % 2).astype(bool), cutoffs=a.astype(np.int64) + 1)
    assert np.array_equal(pgb["iKey"], a % 2)
    assert pgb["iFirstKey"][1].tolist() == [0, 1, 1, 2, 2, 3, 3, 4, 4, 5]


def test_reindex_groups_equals_global_grouping():
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 500, 20_000)
    cuts = np.array([1, 1, 4_000, 9_500, 20_000], dtype=np.int64)
    new, U = rt_glue.hstack_reindex(orc.MultiKeyGroupBy32, orc.ReIndexGroups, keys, cuts)
    gik, _, gU = orc.MultiKeyGroupBy32([keys], 0, None, 2)
    assert U == gU and np.array_equal(new, gik)
    # a 0 (filtered row) stays 0; rows past the last cutoff come back 0
    out = orc.ReIndexGroups(np.array([2, 0, 1, 1, 2], np.int32), np.array([7, 9, 4], np.int32), [2, 3], [3, 4])
    assert out.tolist() == [9, 0, 7, 4, 0]


def test_makeifirst_kat():
    # tests/test_grouping.py:210-296
    for dt in [np.int8, np.int16, np.int32, np.int64]:
        arr = np.ones(174_763, dtype=dt)
        f = orc.MakeiFirst(arr, 2)
        inv = orc.invalid_for(f.dtype)
        assert f[1] == 0 and f[0] == inv and f[2] == inv
        l = orc.MakeiFirst(arr, 2, None, 1)
        assert l[1] == len(arr) - 1 and l[0] == inv and l[2] == inv
    ik = np.array([1, 1, 0, 2, 1], dtype=np.int8)  # Categorical(a,a,b,c,a).set_valid(f)
    f = orc.MakeiFirst(ik, 2)
    assert f[1:].tolist() == [0, 3] and f[0] == orc.invalid_for(f.dtype)
    assert orc.MakeiFirst(ik, 2, None, 1)[1:].tolist() == [4, 3]
    ik = np.array([1, 1, 2, 3, 1], dtype=np.int8)
    flt = np.array([True, True, False, True, True])
    f = orc.MakeiFirst(ik, 3, flt)
    assert f[1] == 0 and f[-1] == 3 and f[2] == orc.invalid_for(f.dtype)
    l = orc.MakeiFirst(ik, 3, flt, 1)
    assert l[1] == 4 and l[-1] == 3


def test_lexsort_matches_numpy_and_hash_equals_lex():
    # tests/test_lexsort.py:21-42, 105-256
    rng = np.random.default_rng(7)
    n = 100_000
    keys = [
        rng.integers(0, 120, n, dtype=np.int8),
        rng.integers(0, 32000, n, dtype=np.int16),
        rng.integers(0, n, n, dtype=np.int32),
        rng.random(n),
        rng.choice(["AAPL₀", "AMZN₂", "IBM₁"], n),
    ]
    assert np.array_equal(orc.LexSort32(keys), np.lexsort(keys))
    arr = rng.choice([np.nan, 1.11, 2.22, 3.33], 50)
    arr[:4] = [1.11, 2.22, 3.33, np.nan]
    gh, gl = orc.groupbyhash(arr, pack=True), orc.groupbylex(arr)
    for k in ("iKey", "iFirstKey", "unique_count", "iGroup", "iFirstGroup", "nCountGroup"):
        assert np.array_equal(gh[k], gl[k]), k
    vn = rng.integers(0, 3, 10_000)
    vs = rng.choice(["a", "b", "c"], 10_000)
    vn[:9] = [0, 0, 0, 1, 1, 1, 2, 2, 2]
    vs[:9] = ["a", "b", "c"] * 3
    gh, gl = orc.groupbyhash([vn, vs], pack=True), orc.groupbylex([vn, vs])
    for k in ("iKey", "iFirstKey", "unique_count", "iGroup", "iFirstGroup", "nCountGroup"):
        assert np.array_equal(gh[k], gl[k]), k
    perm = orc.LexSort32([vn])
    assert np.array_equal(vn[perm][orc.ReverseShuffle(perm)], vn)


def test_lex_all_unique_filter_and_reverse_shuffle():
    # tests/test_lexsort.py:258-283 (all unique: iFirstKey == iGroup == np.lexsort, for int / float / S / U keys),
    # :311-324 (lex grouping under a filter describes the same groups as hash grouping), :325-332 (ReverseShuffle)
    import rt_glue
    rng = np.random.default_rng(4)
    arr = rng.choice(100_000, 50_000, replace=False)
    for dt in (np.int32, np.uint32, np.int64, np.uint64, np.float32, np.float64, "S", "U"):
        a = arr.astype(dt)
        sortidx = np.lexsort([a])
        gbh = rt_glue.groupbyhash(orc, arr, pack=True)
        gbl = rt_glue.groupbylex(orc, a)
        assert np.array_equal(gbl["iFirstKey"], sortidx) and np.array_equal(gbl["iGroup"], sortidx)
        assert gbl["unique_count"] == gbh["unique_count"]
    keys = rng.choice(np.array(["a", "b", "c"]), 20)
    f = np.arange(20) % 2 == 1
    gl, gh = rt_glue.groupbylex(orc, keys, filter=f), rt_glue.groupbyhash(orc, keys, filter=f)
    assert gl["unique_count"] == gh["unique_count"]
    ul, uh = keys[gl["iFirstKey"]], keys[gh["iFirstKey"]]
    exp_l = np.where(gl["iKey"] > 0, np.r_[["Filtered"], ul][gl["iKey"]], "Filtered")
    exp_h = np.where(gh["iKey"] > 0, np.r_[["Filtered"], uh][gh["iKey"]], "Filtered")
    assert np.array_equal(exp_l, exp_h)  # the expanded arrays agree although the numbering differs
    values = rng.integers(1, 7, 300_000)
    sorted_idx = orc.LexSort32([values])
    rev = orc.ReverseShuffle(sorted_idx)
    assert np.array_equal(values[sorted_idx][rev], values)


def test_projections_455654_groups():
    # tests/test_groupby.py:419-459: 1M rows, 4 keys (int, str, str, int), seed 1234 -> exactly 455 654 groups
    n, num_symbols = 1_000_000, 450
    dates = ["20180602", "20180603", "20180604", "20180605", "20180606"]
    exch = np.array(["EXCH1", "EXCH2", "EXCH3"])
    np.random.seed(1234)
    sym = np.random.randint(0, num_symbols, size=n)
    ex = exch[np.random.randint(0, exch.shape[0], size=n)]
    i = np.arange(n)
    td = np.array(dates)[(i * len(dates) / n).astype(int)]
    time_2500 = ((i % (n // len(dates))) / 2500).astype(int)
    r = orc.groupbyhash([sym, ex.astype("S"), td.astype("S"), time_2500])
    assert r["unique_count"] == 455654
    assert (np.diff(r["iFirstKey"]) > 0).all()
    cnt = orc.BinCount(r["iKey"], r["unique_count"] + 1)
    assert cnt[0] == 0 and cnt.sum() == n


def test_ismember_reference_kats():
    # tests/test_ismember.py:14-70 (numbers), :313-322 (strings of different itemsize), :288-300 (index downcast)
    for dt in "bhilqBHILQfd":
        a, b = np.array([1, 2, 3, 4, 5], dtype=dt), np.array([1, 2], dtype=dt)
        m, idx = orc.IsMember32(a, b)
        assert m.tolist() == [True, True, False, False, False] and idx[:2].tolist() == [0, 1]
        assert idx.dtype == np.int8 and (idx[2:] == np.iinfo(np.int8).min).all()
    a = np.array([b"b", b"a", b"a", b"Inv", b"c", b"a", b"b"], dtype="S")
    b = np.array([b"a", b"b", b"c"], dtype="S")
    m, idx = orc.IsMember32(a, b)
    assert m.tolist() == [True, True, True, False, True, True, True]
    assert idx.tolist() == [1, 0, 0, -128, 2, 0, 1] and idx.dtype == np.int8
    for sz, dt in ((5, np.int8), (29_000, np.int16), (60_000, np.int32)):
        _, idx = orc.IsMember32(np.array([b"a", b"b"]), np.full(sz, b"a"))
        assert idx.dtype == dt and idx[0] == 0
    m, idx = orc.IsMember32(np.array([np.nan, 1.0]), np.array([np.nan, 1.0]))
    assert m.tolist() == [False, True] and idx[1] == 1
    m, idx = orc.MultiKeyIsMember32(([np.array([1, 2, 1]), np.array([b"x", b"y", b"z"])],),
                                   ([np.array([1, 1]), np.array([b"z", b"x"])],))
    assert m.tolist() == [True, False, True] and idx[[0, 2]].tolist() == [1, 0]


def test_cutoffs_on_lexsort_pack_are_per_partition_calls_laid_end_to_end():
    """UNPINNED by the reference's tests; DERIVED from tests/test_pgroupby.py:7-41, which pins that a `cutoffs` call on
    the stacked array equals the separate calls on each part (iKey per part, first rows relative to the part,
    cumulative unique cutoffs).  The sort / pack entry points follow the same rule."""
    rng = np.random.default_rng(4)
    arr, arr2 = rng.integers(0, 5, 15), rng.integers(0, 5, 15)
    f1, f2 = np.round(rng.random(15) * 3), np.round(rng.random(15) * 3)
    arr3, f3 = np.hstack([arr, arr2]), np.hstack([f1, f2])
    cut = np.array([15, 30], dtype=np.int64)
    ig = orc.LexSort32([f3, arr3], cutoffs=cut)
    np.testing.assert_array_equal(ig[:15], np.lexsort([f1, arr]))
    np.testing.assert_array_equal(ig[15:], np.lexsort([f2, arr2]))
    ik, ifk, cnt, ucut = orc.GroupFromLexSort(ig, [arr3, f3], cutoffs=cut)
    ik1, ifk1, cnt1 = orc.GroupFromLexSort(ig[:15], [arr, f1])
    ik2, ifk2, cnt2 = orc.GroupFromLexSort(ig[15:], [arr2, f2])
    np.testing.assert_array_equal(ik, np.r_[ik1, ik2])
    np.testing.assert_array_equal(ifk, np.r_[ifk1, ifk2])
    np.testing.assert_array_equal(cnt, np.r_[0, cnt1[1:], cnt2[1:]])
    assert ucut.tolist() == [len(ifk1), len(ifk1) + len(ifk2)]
    # the partitions found by sorting are those the hash route finds (test_pgroupby's own cross-check, multikey form)
    _, (_, hucut), hu = orc.MultiKeyGroupBy32([arr3, f3], cutoffs=cut)
    assert ucut.tolist() == hucut.tolist() and hu == len(ifk)
    # groupbylex carries the fourth result as nUniqueCutoffs and its Python glue still works on the stacked arrays
    hk, _, hu = orc.MultiKeyGroupBy32([arr3], cutoffs=cut)
    g, fg, c = orc.GroupByPack32(hk, None, hu + 1, cutoffs=cut)
    c1, g1, fg1 = orc.BinCount(hk[:15], hk[:15].max() + 1, pack=True)
    c2, g2, fg2 = orc.BinCount(hk[15:], hk[15:].max() + 1, pack=True)
    np.testing.assert_array_equal(g, np.r_[g1, g2])
    np.testing.assert_array_equal(fg, np.r_[fg1, fg2])
    np.testing.assert_array_equal(c, np.r_[c1, c2])
    with pytest.raises(ValueError):
        orc.LexSort32([arr3], cutoffs=np.array([15, 31], dtype=np.int64))


def _cat(values, base_index, cats=None):
    """(codes, sorted uniques) of an ordered Categorical in the given base index."""
    values = np.asarray(values)
    cats = np.unique(values) if cats is None else np.asarray(cats)
    return (np.searchsorted(cats, values) + base_index).astype(np.int8), cats


def test_ismember_categorical_fixup_kats():
    """tests/test_ismember.py:162-195: ismember(c, d) on Categoricals == ismember on their expanded strings, for both
    base indices on both sides; :491-505: filtered and sentinel codes match nothing."""
    rng = np.random.default_rng(0)
    for ba in (0, 1):
        for bb in (0, 1):
            cs = rng.choice(np.array(["a", "b", "c", "d", "e", "f"]), 15)
            ds = rng.choice(np.array(["a", "b", "c"]), 10)
            (c, ccats), (d, dcats) = _cat(cs, ba), _cat(ds, bb)
            for (x, xc, xb, xs), (y, yc, yb, ys) in (((c, ccats, ba, cs), (d, dcats, bb, ds)), ((d, dcats, bb, ds), (c, ccats, ba, cs))):
                _, on_unique = orc.IsMember32(xc, yc)
                b, f = orc.IsMemberCategoricalFixup(x, y, on_unique.astype(np.int32), len(yc), xb, yb)
                bs, fs = orc.IsMember32(xs, ys)
                np.testing.assert_array_equal(b, bs)
                np.testing.assert_array_equal(f[b].astype(np.int8), fs[bs])
                assert (f[~b] == np.iinfo(np.int32).min).all()
    # tests/test_ismember.py:491-505
    c1 = np.array([0, -128, 3], dtype=np.int8)           # ["a", "b", "c"] with row 0 filtered and row 1 set to the sentinel
    c2 = np.array([1, 2, -128, 0, 3], dtype=np.int8)     # ["a", "b", "c", "d", "c"] with row 3 filtered and row 2 the sentinel
    is_in, idx = orc.IsMemberCategoricalFixup(c1, c2, np.array([0, 1, 2], dtype=np.int32), 4, 1, 1)
    assert is_in.tolist() == [False, False, True]
    assert idx.tolist() == [np.iinfo(np.int32).min, np.iinfo(np.int32).min, 4]


# ---- 8f.4: join-index construction (oracle/merge_oracle.py restates rt_merge.py; pinned to pandas.merge) --------------
def _pandas_join(lk, rk, how, keep):
    import pandas as pd

    l = pd.DataFrame({f"k{i}": c for i, c in enumerate(lk)})
    r = pd.DataFrame({f"k{i}": c for i, c in enumerate(rk)})
    l["lid"], r["rid"] = np.arange(len(l)), np.arange(len(r))
    on = [c for c in l.columns if c.startswith("k")]
    if keep[0]:
        l = l.drop_duplicates(on, keep=keep[0])
    if keep[1]:
        r = r.drop_duplicates(on, keep=keep[1])
    m = pd.merge(l, r, on=on, how=how)
    return m["lid"].to_numpy(dtype=np.float64), m["rid"].to_numpy(dtype=np.float64)


def _as_float(idx, n):
    from oracle import merge_oracle as mo

    if idx is None:
        return np.arange(n, dtype=np.float64)
    out = np.asarray(idx).astype(np.float64)
    out[np.asarray(idx) == mo.INV] = np.nan
    return out


@pytest.mark.parametrize("how", ["left", "inner", "right"])
@pytest.mark.parametrize("keep", [(None, None), (None, "first"), ("last", None), ("first", "last")])
def test_merge_indices_row_order_equals_pandas_merge(how, keep):
    """rt_merge.py:2107-2113: left / right / inner joins preserve the key order of the left (right) side; matching rows of
    the other side come in their own row order -- which is pandas.merge's contract too."""
    from oracle import merge_oracle as mo

    rng = np.random.default_rng(hash((how, keep)) % 2**32)
    for nl, nr, span in ((0, 5, 3), (7, 0, 3), (50, 40, 12), (300, 500, 40), (200, 30, 500)):
        for cols in (1, 2):
            lk = [rng.integers(0, span, nl) for _ in range(cols)]
            rk = [rng.integers(0, span, nr) for _ in range(cols)]
            if cols == 2:
                lk[1] = np.array(["x", "yy", "z"])[lk[1] % 3]
                rk[1] = np.array(["x", "yy", "z"])[rk[1] % 3]
            li, ri, _ = mo.merge_indices(lk, rk, how, keep)
            pl, pr = _pandas_join(lk, rk, how, keep)
            n_out = len(pl)
            gl = _as_float(li, nl)
            gr = _as_float(ri, nr)
            assert len(gl) == len(gr) == n_out, (how, keep, nl, nr)
            driving_keep = keep[1] if how == "right" else keep[0]
            if driving_keep == "last":
                # riptable walks the driving side's groups in first-appearance order and takes each group's LAST row
                # (rt_merge.py:785-789: ikey[ilastkey]); pandas keeps the surviving rows in row order.  Same pairs.
                key = lambda a, b: sorted(zip(np.nan_to_num(a, nan=-1).tolist(), np.nan_to_num(b, nan=-1).tolist()))  # noqa: E731
                assert key(gl, gr) == key(pl, pr)
                drv, dk = (gr, rk) if how == "right" else (gl, lk)
                rows = drv[~np.isnan(drv)].astype(np.int64)
                bins, _, _ = orc.MultiKeyGroupBy32([np.asarray(c) for c in dk]) if len(dk[0]) else (np.zeros(0, np.int32), None, 0)
                order = bins[rows].tolist()  # first-appearance number of each surviving row's key
                assert order == sorted(order)
            else:
                np.testing.assert_array_equal(gl, pl)
                np.testing.assert_array_equal(gr, pr)


def test_merge_indices_outer_and_null_keys():
    """Outer join (rt_merge.py:1404-1588): the left-join rows first, then the rows of right whose key is not in left, in
    right row order.  NaN keys never match (SQL JOIN semantics, rt_merge.py:1968-1981) but their rows are kept."""
    from oracle import merge_oracle as mo

    lk = [np.array([3.0, np.nan, 1.0, 3.0, 7.0])]
    rk = [np.array([1.0, 1.0, np.nan, 9.0, 3.0, 8.0])]
    li, ri, k = mo.merge_indices(lk, rk, "outer")
    assert k == 3
    assert _as_float(li, 5).tolist()[:6] == [0.0, 1.0, 2.0, 2.0, 3.0, 4.0] and np.isnan(_as_float(li, 5)[6:]).all()
    r = _as_float(ri, 6)
    assert r[[0, 2, 3, 4]].tolist() == [4.0, 0.0, 1.0, 4.0] and np.isnan(r[[1, 5]]).all() and r[6:].tolist() == [2.0, 3.0, 5.0]
    # left join: the left NaN row stays (and matches nothing); inner: it goes
    li, ri, _ = mo.merge_indices(lk, rk, "left")
    assert _as_float(li, 5).tolist() == [0.0, 1.0, 2.0, 2.0, 3.0, 4.0]
    li, ri, _ = mo.merge_indices(lk, rk, "inner")
    assert li.tolist() == [0, 2, 2, 3] and ri.tolist() == [4, 0, 1, 4]
    # unique right keys: the left columns are used as they are (rt_merge.py:1060-1095)
    li, ri, _ = mo.merge_indices([np.array([5, 6, 5])], [np.array([6, 5, 9])], "left")
    assert li is None and ri.tolist() == [1, 0, 1]
    # outer with keep on the right: one row per right-only key, all pairs as a set equal pandas' outer join
    rng = np.random.default_rng(1)
    lk, rk = [rng.integers(0, 30, 80)], [rng.integers(10, 50, 90)]
    for keep in ((None, None), (None, "first"), (None, "last")):
        li, ri, k = mo.merge_indices(lk, rk, "outer", keep)
        pl, pr = _pandas_join(lk, rk, "outer", keep)
        got = sorted(zip(np.nan_to_num(_as_float(li, 80), nan=-1).tolist(), np.nan_to_num(_as_float(ri, 90), nan=-1).tolist()))
        exp = sorted(zip(np.nan_to_num(pl, nan=-1).tolist(), np.nan_to_num(pr, nan=-1).tolist()))
        assert got == exp
        tail = _as_float(ri, 90)[len(ri) - k:]
        assert (np.diff(tail) > 0).all() and np.isnan(_as_float(li, 80)[len(li) - k:]).all()
    with pytest.raises(ValueError):
        mo.merge_indices(lk, rk, "outer", ("first", None))
