"""GPU parity: rc.MultiKeyGroupBy32 / GroupByAll32 / BinCount / CombineFilter vs the CPU oracle.
Integer outputs bit-exact; float reductions within 1e-12 (f64) / 1e-5 (f32) relative."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import riptide_oracle as orc


@pytest.fixture(scope="module")
def rc():
    import riptable_b200.rc as rc

    return rc


def _check_gb(rc, keys, filter=None, cutoffs=None, hint=0):
    ik, ifk, u = rc.MultiKeyGroupBy32(keys, hint, filter, 2, cutoffs=cutoffs)
    ok, ofk, ou = orc.MultiKeyGroupBy32(keys, hint, filter, 2, cutoffs=cutoffs)
    assert u == ou
    assert ik.dtype == np.int32
    np.testing.assert_array_equal(ik, ok)
    if cutoffs is None:
        assert ifk.dtype == np.int32
        np.testing.assert_array_equal(ifk, ofk)
    else:
        np.testing.assert_array_equal(ifk[0], ofk[0])
        np.testing.assert_array_equal(ifk[1], ofk[1])
    return ik, u


@pytest.mark.parametrize("dt", [np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64,
                                np.float32, np.float64, np.bool_])
@pytest.mark.parametrize("n", [1, 5, 1000, 100_003])
def test_single_key_dtypes(rc, dt, n):
    rng = np.random.default_rng(n)
    if dt == np.bool_:
        k = rng.integers(0, 2, n).astype(dt)
    else:
        k = rng.integers(0, 100, n).astype(dt)
    _check_gb(rc, [k])
    _check_gb(rc, [k], filter=rng.random(n) < 0.6)


def test_docstring_vectors(rc):
    # rt_numpy.py:2013-2062
    np.random.seed(12345)
    c = np.random.randint(0, 8000, 2_000_000)
    d = np.random.randint(0, 8000, 2_000_000)
    ik, ifk, u = rc.MultiKeyGroupBy32([c], 0, None, 2)
    assert u == 8000 and ik[:3].tolist() == [1, 2, 3] and ik[-3:].tolist() == [6061, 7889, 3002]
    assert ifk[-3:].tolist() == [67072, 67697, 68250]
    ik, ifk, u = rc.MultiKeyGroupBy32([c], 0, (c % 3).astype(bool), 2)
    assert u == 5333 and ik[-3:].tolist() == [0, 5250, 1973] and ifk[-3:].tolist() == [54422, 58655, 68250]
    ik, ifk, u = rc.MultiKeyGroupBy32([c, d], 0, None, 2)
    assert u == 1968856 and ik[-3:].tolist() == [1968854, 1968855, 1968856]
    assert ifk[-3:].tolist() == [1999997, 1999998, 1999999]
    _check_gb(rc, [c, d])


def test_special_keys(rc):
    # -1 collides with the table's empty marker; NaNs group bytewise; -0.0 != +0.0 bytewise
    k = np.array([-1, 5, -1, 0, 5, -1, np.iinfo(np.int64).min], dtype=np.int64)
    _check_gb(rc, [k])
    k = np.tile(np.array([-1, -1, 3], dtype=np.int64), 100_000)
    _check_gb(rc, [k])
    f = np.array([np.nan, 1.0, np.nan, -0.0, 0.0, 1.0, np.inf, -np.inf, np.nan])
    _check_gb(rc, [f])
    _check_gb(rc, [f.astype(np.float32)])


@pytest.mark.parametrize("n,u", [(3_000_000, 2_500_000), (2_000_000, 2_000_000), (5_000_000, 1000)])
def test_high_cardinality_and_growth(rc, n, u):
    rng = np.random.default_rng(u)
    vals = rng.integers(-(2**62), 2**62, u, dtype=np.int64)
    k = vals[rng.integers(0, u, n)] if u < n else rng.permutation(vals)
    _check_gb(rc, [k])


def test_sorted_and_late_keys(rc):
    # discovery rate that never decays (sorted keys) and a burst of new keys late in the array
    k = np.repeat(np.arange(300_000, dtype=np.int64), 10)
    _check_gb(rc, [k])
    rng = np.random.default_rng(3)
    k = np.concatenate([rng.integers(0, 50, 2_500_000), rng.integers(1000, 400_000, 600_000)]).astype(np.int32)
    _check_gb(rc, [k])


def test_multikey_and_strings(rc):
    rng = np.random.default_rng(11)
    n = 200_000
    a = rng.integers(0, 50, n)
    s = rng.choice(np.array(["alpha", "be", "gamma_long_string", ""], dtype="S17"), n)
    u = rng.choice(np.array(["x", "yy", "zzz"], dtype="U3"), n)
    i4 = rng.integers(0, 5, n).astype(np.int32)
    _check_gb(rc, [s])
    _check_gb(rc, [u])
    _check_gb(rc, [a, s, i4])
    _check_gb(rc, [a, s, i4], filter=rng.random(n) < 0.5)
    _check_gb(rc, [i4, i4.astype(np.int8), u])
    s16 = rng.choice(np.array([b"ABCDEFGHIJKLMNOP", b"ABCDEFGHIJKLMNOQ", b"ZBCDEFGHIJKLMNOP"], dtype="S16"), n)
    _check_gb(rc, [a.astype(np.int64), s16, i4])


def test_small_rows_composed_into_one_key(rc):
    # key rows of at most 8 bytes (several small columns, or one 3/5/6/7-byte column) run as one composed 64-bit key
    rng = np.random.default_rng(41)
    n = 2_400_007
    a4, b4 = rng.integers(-50, 50, n).astype(np.int32), rng.integers(0, 300, n).astype(np.int32)
    a2, a1 = rng.integers(-5, 5, n).astype(np.int16), rng.integers(0, 3, n).astype(np.int8)
    f4 = rng.choice(np.array([0.0, -0.0, 1.5, np.nan], dtype=np.float32), n)
    s3 = rng.choice(np.array([b"", b"a", b"ab", b"abc", b"zzz"], dtype="S3"), n)
    s5 = rng.choice(rng.integers(65, 91, 700 * 5, dtype=np.uint8).view("S5"), n)
    filt = rng.random(n) < 0.7
    _check_gb(rc, [a4, b4])
    _check_gb(rc, [a4, b4], filter=filt)
    _check_gb(rc, [a2, a4, a1])
    _check_gb(rc, [f4, a2])
    _check_gb(rc, [s3, a4])
    _check_gb(rc, [s5])
    _check_gb(rc, [s3], filter=filt)
    _check_gb(rc, [np.full(n, -1, np.int32), np.full(n, -1, np.int32)])  # composed key == the table's empty marker
    _check_gb(rc, [a4[:777], b4[:777]])


def test_packed_rows_in_shared_memory(rc):
    # settled string / multi-key tables small enough for shared memory (gb_round_packed_smem): rounds past the first 1 Mi
    # rows, with a filter, and with keys that first appear after the table had settled (they take the global fallback)
    rng = np.random.default_rng(31)
    n = 3_300_001
    s10 = rng.choice(rng.integers(65, 91, 300 * 10, dtype=np.uint8).view("S10"), n)
    u8 = rng.choice(rng.integers(65, 91, 200 * 8, dtype=np.int32).view("<U8"), n)
    i8 = rng.integers(0, 40, n).astype(np.int64)
    i4 = rng.integers(0, 5, n).astype(np.int32)
    s16 = rng.choice(rng.integers(65, 91, 50 * 16, dtype=np.uint8).view("S16"), n)
    filt = rng.random(n) < 0.6
    _check_gb(rc, [s10])
    _check_gb(rc, [u8], filter=filt)
    _check_gb(rc, [s16])
    _check_gb(rc, [i8, s16, i4])
    _check_gb(rc, [s10, i4], filter=filt)
    late = s10.copy()
    late[2_600_000:] = rng.choice(rng.integers(97, 123, 700 * 10, dtype=np.uint8).view("S10"), n - 2_600_000)
    _check_gb(rc, [late])
    many = rng.choice(rng.integers(65, 91, 9_000 * 16, dtype=np.uint8).view("S16"), n)  # 9 000 keys of 16 bytes: 144 KB
    _check_gb(rc, [many])


def test_packed_rows_snapshot_table(rc):
    # 16-byte rows with more keys than shared memory holds: settled rounds probe the inline-key snapshot; keys that
    # first appear later take the exact path and are added to the snapshot for the following rounds
    rng = np.random.default_rng(53)
    n = 6_000_011
    a8 = rng.integers(0, 3000, n).astype(np.int64) * 1_000_003
    b4 = rng.integers(0, 40, n).astype(np.int32)
    _check_gb(rc, [a8, b4])  # 120 000 pairs
    _check_gb(rc, [a8, b4], filter=rng.random(n) < 0.5)
    s12 = rng.choice(rng.integers(65, 91, 40_000 * 12, dtype=np.uint8).view("S12"), n)
    _check_gb(rc, [s12])
    late = a8.copy()
    late[4_500_000:] += rng.integers(0, 2, n - 4_500_000) * 17  # new pairs appear after the table had settled
    _check_gb(rc, [late, b4])
    s8 = rng.choice(rng.integers(65, 91, 5000 * 8, dtype=np.uint8).view("S8"), n)
    _check_gb(rc, [s8, b4])
    # 32-byte rows: U8 keys, and (int64, S16, int32) = 28 bytes
    u8 = rng.choice(rng.integers(65, 91, 12_000 * 8, dtype=np.int32).view("<U8"), n)
    _check_gb(rc, [u8])
    s16 = rng.choice(rng.integers(65, 91, 300 * 16, dtype=np.uint8).view("S16"), n)
    _check_gb(rc, [a8, s16, b4 % 3], filter=rng.random(n) < 0.8)


def test_projections_fixture(rc):
    # tests/test_groupby.py:419-459 -> exactly 455 654 groups
    n, num_symbols = 1_000_000, 450
    dates = ["20180602", "20180603", "20180604", "20180605", "20180606"]
    exch = np.array(["EXCH1", "EXCH2", "EXCH3"])
    np.random.seed(1234)
    sym = np.random.randint(0, num_symbols, size=n)
    ex = exch[np.random.randint(0, exch.shape[0], size=n)]
    i = np.arange(n)
    td = np.array(dates)[(i * len(dates) / n).astype(int)]
    t25 = ((i % (n // len(dates))) / 2500).astype(int)
    ik, u = _check_gb(rc, [sym, ex.astype("S"), td.astype("S"), t25])
    assert u == 455654


def test_host_mode_large_pageable_inputs(rc):
    # NumPy inputs above 8 MB travel through the staged page-locked ring (rc._h2d_staged); sizes straddle chunk edges
    rng = np.random.default_rng(77)
    for n in (1 << 20, (1 << 21) + 12_345, 3_000_001):
        key = rng.integers(0, 5000, n).astype(np.int64) * 104_729
        val = rng.random(n)
        ik, ifk, u = rc.MultiKeyGroupBy32([key], 0, None, 2)
        eik, eifk, eu = orc.MultiKeyGroupBy32([key], 0, None, 2)
        assert u == eu
        np.testing.assert_array_equal(ik, eik)
        np.testing.assert_array_equal(ifk, eifk)
        got = rc.GroupByAll32([val], ik, u, [0], [1], [u + 1], 0)[0]
        np.testing.assert_allclose(got[1:], np.bincount(eik, weights=val, minlength=eu + 1)[1:], rtol=1e-12)
        d = rc._to_dev(key)
        assert (d.torch().cpu().numpy() == key).all()


def test_ismember_categorical(rc):
    rng = np.random.default_rng(21)
    for na, nb in ((50_000, 80), (20_000, 40_000), (10, 0), (0, 10)):
        b = rng.permutation(max(nb, 1) * 2)[:nb].astype(np.int32)
        a = rng.integers(-1, max(nb, 1) * 2, na).astype(np.int32)
        ef, ei = orc.IsMemberCategorical(a, b)
        gf, gi = rc.IsMemberCategorical(a, b)
        assert gi.dtype == ei.dtype
        np.testing.assert_array_equal(gf, ef)
        np.testing.assert_array_equal(gi, ei)
    keys = np.array(["pear", "apple", "fig", "apple", "pear", "kiwi"])
    ik, ifk, U = rc.MultiKeyGroupBy32([keys], 0, np.array([True, True, True, True, True, False]), 2)
    sortidx = np.argsort(keys[ifk], kind="stable")
    _, new = rc.IsMemberCategorical(ik.astype(np.int32) - 1, sortidx.astype(np.int32))
    assert new.tolist() == [3, 1, 2, 1, 3, 0]


def test_reindex_groups(rc):
    # 8f.3: the hstack_groupings recipe (rt_grouping.py:277-413) must land on the global grouping
    import rt_glue
    rng = np.random.default_rng(3)
    keys = rng.integers(0, 50_000, 700_001)
    cuts = np.array([1, 1, 40_000, 95_000, 500_003, 700_001], dtype=np.int64)
    new, U = rt_glue.hstack_reindex(rc.MultiKeyGroupBy32, rc.ReIndexGroups, keys, cuts)
    gik, _, gU = orc.MultiKeyGroupBy32([keys], 0, None, 2)
    assert U == gU and np.array_equal(new, gik)
    ik, (ifk, ucuts), _ = orc.MultiKeyGroupBy32([keys], 0, None, 2, cutoffs=cuts)
    uik = rng.integers(1, 1000, int(ucuts[-1])).astype(np.int32)
    ik = ik.copy()
    ik[rng.random(len(ik)) < 0.1] = 0
    for dt in (np.int32, np.int64):
        got = rc.ReIndexGroups(ik.astype(dt), uik, ucuts, cuts)
        exp = orc.ReIndexGroups(ik.astype(dt), uik, ucuts, cuts)
        assert got.dtype == exp.dtype and np.array_equal(got, exp)
    out = rc.ReIndexGroups(np.array([2, 0, 1, 1, 2], np.int32), np.array([7, 9, 4], np.int32), [2, 3], [3, 4])
    assert out.tolist() == [9, 0, 7, 4, 0]


def test_cutoffs(rc):
    # tests/test_pgroupby.py:7-86
    rng = np.random.default_rng(0)
    arr3 = rng.integers(0, 5, 30)
    _check_gb(rc, [arr3], cutoffs=np.array([15, 30], dtype=np.int64))
    _check_gb(rc, [arr3], cutoffs=np.array([30], dtype=np.int64))
    a = np.arange(10)
    _check_gb(rc, [np.ones(10)], cutoffs=a.astype(np.int64) + 1)
    _check_gb(rc, [np.ones(10)], filter=(a % 2).astype(bool), cutoffs=a.astype(np.int64) + 1)
    big = rng.integers(0, 1000, 300_000)
    _check_gb(rc, [big], cutoffs=np.array([1, 1, 100_000, 250_000, 300_000], dtype=np.int64))
    _check_gb(rc, [big, big % 7], filter=rng.random(300_000) < 0.7, cutoffs=np.array([123_457, 300_000], dtype=np.int64))


def test_errors_and_empty(rc):
    ik, ifk, u = rc.MultiKeyGroupBy32([np.zeros(0, np.int64)], 0, None, 2)
    assert u == 0 and len(ik) == 0 and len(ifk) == 0
    with pytest.raises(ValueError):
        rc.MultiKeyGroupBy32([np.zeros(3), np.zeros(4)], 0, None, 2)
    with pytest.raises(ValueError):
        rc.MultiKeyGroupBy32([], 0, None, 2)


VAL_DTYPES = [np.bool_, np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32, np.int64, np.uint64, np.float32,
              np.float64]
FUNCS = [0, 1, 2, 3, 4, 5, 50, 51, 52, 53, 54, 55]


def _values(rng, dt, n, invalid_frac):
    dt = np.dtype(dt)
    if dt.kind == "b":
        return rng.integers(0, 2, n).astype(dt)
    if dt.kind == "f":
        v = (rng.random(n) * 2e4 - 1e4).astype(dt)
    elif dt.kind == "i":
        v = rng.integers(-100 if dt.itemsize == 1 else -10_000, 100 if dt.itemsize == 1 else 10_000, n).astype(dt)
    else:
        v = rng.integers(0, 200 if dt.itemsize == 1 else 20_000, n).astype(dt)
    if invalid_frac > 0:
        v[rng.random(n) < invalid_frac] = orc.invalid_for(dt)
    return v


def _cmp_reduce(got, exp, vdt):
    assert got.dtype == exp.dtype, (got.dtype, exp.dtype)
    if exp.dtype.kind == "f":
        rtol = 1e-5 if np.dtype(vdt) == np.float32 else 1e-12
        # relative to the magnitude of the data, not of a possibly cancelling result
        np.testing.assert_allclose(got, exp, rtol=rtol, atol=rtol * 2e4 * 50, equal_nan=True)
    else:
        np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("kdt", [np.int8, np.int16, np.int32, np.int64])
@pytest.mark.parametrize("nb", [3, 90, 20_000])
def test_reduce_all_ops_all_dtypes(rc, kdt, nb):
    if nb > np.iinfo(kdt).max:
        pytest.skip("bins do not fit the key dtype")
    rng = np.random.default_rng(nb)
    n = 150_001
    ikey = rng.integers(0, nb + 1, n).astype(kdt)
    ikey[ikey == 2] = 1  # bin 2 stays empty
    for vdt in VAL_DTYPES:
        for frac in (0.0, 0.2):
            v = _values(rng, vdt, n, frac)
            for fn in FUNCS:
                for lo in (1, 0):
                    got = rc.GroupByAll32([v], ikey, nb, [fn], [lo], [nb + 1], 0)[0]
                    exp = orc.GroupByAll32([v], ikey, nb, [fn], [lo], [nb + 1], 0)[0]
                    _cmp_reduce(got[lo:], exp[lo:], vdt)
                    if lo == 1 and fn in (0, 50):
                        assert got[0] == 0


def test_reduce_unsupported_and_kat(rc):
    ikey = np.array([2, 1, 2, 1, 2], dtype=np.int8)
    s = np.array([b"3", b"4", b"5", b"1", b"2"])
    v = np.array([3, 4, 5, 1, 2], dtype=np.int32)
    out = rc.GroupByAll32([s, v], ikey, 2, [0, 0], [1, 1], [3, 3], 0)
    assert out[0] is None and out[1][1:].tolist() == [5, 10] and out[1].dtype == np.int64
    # all-invalid and extreme values
    x = np.array([np.iinfo(np.int64).max] * 3 + [np.iinfo(np.int64).min] * 2, dtype=np.int64)
    ik = np.array([1, 1, 1, 2, 2], dtype=np.int32)
    for fn in (2, 3, 52, 53):
        got = rc.GroupByAll32([x], ik, 3, [fn], [1], [4], 0)[0]
        exp = orc.GroupByAll32([x], ik, 3, [fn], [1], [4], 0)[0]
        np.testing.assert_array_equal(got, exp)
    xu = np.array([0, 0, np.iinfo(np.uint64).max, 5], dtype=np.uint64)
    ik = np.array([1, 1, 2, 2], dtype=np.int32)
    for fn in (2, 3, 52, 53):
        np.testing.assert_array_equal(rc.GroupByAll32([xu], ik, 2, [fn], [1], [3], 0)[0],
                                      orc.GroupByAll32([xu], ik, 2, [fn], [1], [3], 0)[0])


def test_bincount_and_combine_filter(rc):
    rng = np.random.default_rng(5)
    for kdt, nb in [(np.int8, 50), (np.int16, 3000), (np.int32, 70_000), (np.int64, 10)]:
        n = 333_333
        ikey = rng.integers(0, nb, n).astype(kdt)
        got = rc.BinCount(ikey, nb)
        assert got.dtype == np.int32
        np.testing.assert_array_equal(got, np.bincount(ikey, minlength=nb))
        f = rng.random(n) < 0.5
        cf = rc.CombineFilter(ikey, f)
        assert cf.dtype == ikey.dtype
        np.testing.assert_array_equal(cf, np.where(f, ikey, 0))


def test_device_mode_chain(rc):
    import torch

    rng = np.random.default_rng(9)
    n = 1_000_000
    k = rng.integers(0, 5000, n).astype(np.int64) * 7919
    v = rng.random(n)
    kt, vt = torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda()
    ik, ifk, u = rc.MultiKeyGroupBy32([kt], 0, None, 2)
    assert ik.is_cuda and ifk.is_cuda and ik.dtype == torch.int32
    s = rc.GroupByAll32([vt], ik, u, [0], [1], [u + 1], 0)[0]
    assert s.is_cuda
    ok, ofk, ou = orc.MultiKeyGroupBy32([k])
    assert u == ou
    np.testing.assert_array_equal(ik.cpu().numpy(), ok)
    exp = orc.GroupByAll32([v], ok, ou, [0], [1], [ou + 1], 0)[0]
    np.testing.assert_allclose(s.cpu().numpy()[1:], exp[1:], rtol=1e-12)


def test_combine_accum_filters(rc):
    # rt_numpy.py:1533-1540 docstring: key = (arange(20) % 10) + 1, filter = odd rows
    key = ((np.arange(20) % 10) + 1).astype(np.int8)
    f = (np.arange(20) % 2).astype(bool)
    ik, ifk, u = rc.CombineAccum1Filter(key, 10, f)
    assert ik.tolist() == [0, 1, 0, 2, 0, 3, 0, 4, 0, 5, 0, 1, 0, 2, 0, 3, 0, 4, 0, 5]
    assert ik.dtype == np.int8 and ifk.tolist() == [1, 3, 5, 7, 9] and u == 5
    rng = np.random.default_rng(1)
    n = 100_000
    k1 = rng.integers(0, 41, n).astype(np.int16)
    k2 = rng.integers(0, 7, n).astype(np.int8)
    flt = rng.random(n) < 0.6
    for filt in (None, flt):
        gik, gifk, gu = rc.CombineAccum1Filter(k1, 40, filt)
        eik, eifk, eu = orc.CombineAccum1Filter(k1, 40, filt)
        assert gu == eu and gik.dtype == eik.dtype
        np.testing.assert_array_equal(gik, eik)
        np.testing.assert_array_equal(gifk, eifk)
        gik, gcnt = rc.CombineAccum2Filter(k1, k2, 40, 6, filt)
        eik, ecnt = orc.CombineAccum2Filter(k1, k2, 40, 6, filt)
        assert gik.dtype == eik.dtype and gcnt.dtype == np.int32
        np.testing.assert_array_equal(gik, eik)
        np.testing.assert_array_equal(gcnt, ecnt)


def test_groupby_stream_chunks_equal_one_shot(rc):
    # the resumable form: chunked calls == one call on the concatenation (bins keep counting across chunks)
    import torch

    rng = np.random.default_rng(21)
    for dt, n, u, cuts in ((np.int64, 3_000_000, 200_000, [0, 1_000_000, 1_000_128, 2_500_000, 3_000_000]),
                           (np.int32, 500_000, 3_000, [0, 100_000, 500_000]),
                           (np.int16, 200_001, 50, [0, 1, 65, 200_001])):
        vals = rng.integers(np.iinfo(dt).min // 2, np.iinfo(dt).max // 2, u).astype(dt)
        k = vals[rng.integers(0, u, n)]
        k[: n // 2] = vals[rng.integers(0, u // 2 + 1, n // 2)]  # half of the keys only show up late
        filt = rng.random(n) < 0.9
        for f in (None, filt):
            eik, eifk, eu = orc.MultiKeyGroupBy32([k], 0, f, 2)
            st = rc.GroupByStream(max_unique=max(eu, 1024), max_chunk_rows=max(b - a for a, b in zip(cuts, cuts[1:])))
            parts = []
            for a, b in zip(cuts, cuts[1:]):
                kc = torch.from_numpy(k[a:b].copy()).cuda()
                fc = None if f is None else torch.from_numpy(f[a:b].copy()).cuda()
                parts.append(st.push(kc, fc, row_offset=a).cpu().numpy())
            assert st.unique_count == eu
            np.testing.assert_array_equal(np.concatenate(parts), eik)
            np.testing.assert_array_equal(st.ifirstkey.cpu().numpy(), eifk)


def test_ismember_vs_oracle(rc):
    # 8f.1: rc.IsMember32 / rc.MultiKeyIsMember32 (tests/test_ismember.py semantics)
    rng = np.random.default_rng(44)
    for dt in (np.int8, np.uint16, np.int32, np.int64, np.uint64, np.float32, np.float64):
        a = rng.integers(0, 100, 50_000).astype(dt)
        b = rng.integers(50, 120, 777).astype(dt)
        if np.dtype(dt).kind == "f":
            a[::7] = np.nan
            b[::5] = np.nan
        m, idx = rc.IsMember32(a, b)
        em, eidx = orc.IsMember32(a, b)
        assert idx.dtype == eidx.dtype
        np.testing.assert_array_equal(m, em)
        np.testing.assert_array_equal(idx, eidx)
    # the lookup path: large tables, the all-ones key (the table's empty marker), all-NaN b, bool keys, device mode
    for dt, lo, hi in ((np.int64, -3, 200_000), (np.uint64, 0, 90_000), (np.int32, -70_000, 70_000)):
        a = rng.integers(lo, hi, 300_001).astype(dt)
        b = rng.integers(lo, hi, 120_000).astype(dt)
        a[::11] = np.array(-1).astype(dt) if np.dtype(dt).kind == "i" else np.iinfo(dt).max
        for bb in (b, np.r_[b, a[:1].repeat(3)]):
            m, idx = rc.IsMember32(a, bb)
            em, eidx = orc.IsMember32(a, bb)
            assert idx.dtype == eidx.dtype
            np.testing.assert_array_equal(m, em)
            np.testing.assert_array_equal(idx, eidx)
    m, idx = rc.IsMember32(np.array([1.0, np.nan]), np.array([np.nan, np.nan]))
    assert m.tolist() == [False, False]
    m, idx = rc.IsMember32(np.array([True, False, True]), np.array([True]))
    assert m.tolist() == [True, False, True] and idx.tolist() == [0, -128, 0]
    import torch
    ta, tb = torch.arange(10, device="cuda", dtype=torch.int64), torch.tensor([7, 3, 3], device="cuda", dtype=torch.int64)
    m, idx = rc.IsMember32(ta, tb)
    assert m.is_cuda and m.cpu().tolist() == [False, False, False, True, False, False, False, True, False, False]
    assert idx.cpu().tolist()[3] == 1 and idx.cpu().tolist()[7] == 0
    a = np.array([b"b", b"a", b"a", b"Inv", b"c", b"a", b"b"], dtype="S")
    b = np.array([b"a", b"b", b"c"], dtype="S")
    m, idx = rc.IsMember32(a, b)
    assert m.tolist() == [True, True, True, False, True, True, True] and idx.tolist() == [1, 0, 0, -128, 2, 0, 1]
    for sz, dt in ((5, np.int8), (29_000, np.int16), (60_000, np.int32)):
        _, idx = rc.IsMember32(np.array([b"a", b"b"]), np.full(sz, b"a"))
        assert idx.dtype == dt and idx[0] == 0 and idx[1] == np.iinfo(dt).min
    n = 200_000
    a1, a2 = rng.integers(0, 300, n), rng.choice(np.array(["x", "yy", "zzz", ""]), n)
    b1, b2 = rng.integers(100, 400, 5000), rng.choice(np.array(["x", "yy", "w"]), 5000)
    m, idx = rc.MultiKeyIsMember32(([a1, a2],), ([b1, b2],))
    em, eidx = orc.MultiKeyIsMember32(([a1, a2],), ([b1, b2],))
    np.testing.assert_array_equal(m, em)
    np.testing.assert_array_equal(idx, eidx)
    m, idx = rc.IsMember32(np.zeros(0, np.int32), np.arange(3, dtype=np.int32))
    assert len(m) == 0 and len(idx) == 0
    m, idx = rc.IsMember32(np.arange(3, dtype=np.int32), np.zeros(0, np.int32))
    assert not m.any() and (idx == np.iinfo(np.int8).min).all()


# ---- fused grouping + sum (rtb_multikey_groupby_sum32): must equal the two reference calls it stands for ----------
def _check_fused(rc, keys, val, filter=None, func=0, bin_low=1, hint=0):
    ik, ifk, u, s = rc.MultiKeyGroupBySum32(keys, val, hint, filter, 2, func, bin_low)
    ok, ofk, ou = orc.MultiKeyGroupBy32(keys, hint, filter, 2)
    assert u == ou and ik.dtype == np.int32 and ifk.dtype == np.int32
    np.testing.assert_array_equal(ik, ok)
    np.testing.assert_array_equal(ifk, ofk)
    exp = orc.GroupByAll32([val], ok, ou, [func], [bin_low], [ou + 1], 0)[0]
    assert s.dtype == exp.dtype and s.shape == exp.shape
    if exp.dtype.kind == "f":
        rtol = 1e-12 if exp.dtype == np.float64 else 1e-5
        scale = max(1.0, float(np.nanmax(np.abs(val.astype(np.float64)))) if len(val) else 1.0)
        cnt = np.bincount(ok, minlength=ou + 1)
        np.testing.assert_allclose(s[bin_low:], exp[bin_low:], rtol=rtol, atol=rtol * scale * max(1.0, float(cnt.max()) ** 0.5))
    else:
        np.testing.assert_array_equal(s[bin_low:], exp[bin_low:])
    return u


@pytest.mark.parametrize("kdt", [np.int64, np.int32, np.uint16, np.int8, np.float64])
@pytest.mark.parametrize("vdt", [np.float64, np.float32, np.int64, np.int32, np.int16, np.uint8, np.bool_])
@pytest.mark.parametrize("n,u", [(1, 1), (130, 7), (100_003, 900), (3_000_017, 50_000)])
def test_fused_groupby_sum(rc, kdt, vdt, n, u):
    rng = np.random.default_rng(n + u)
    info_hi = 100 if np.dtype(kdt).itemsize == 1 else u
    uv = rng.permutation(np.arange(-(info_hi // 2), info_hi - info_hi // 2))[: min(u, info_hi)]
    if np.dtype(kdt).kind == "u":
        uv = np.abs(uv)
    k = uv[rng.integers(0, len(uv), n)].astype(kdt)
    if np.dtype(vdt).kind == "f":
        v = (rng.random(n) * 2 - 1).astype(vdt)
    elif vdt == np.bool_:
        v = rng.random(n) < 0.5
    else:
        v = rng.integers(0, 100, n).astype(vdt)
    _check_fused(rc, [k], v)
    _check_fused(rc, [k], v, filter=rng.random(n) < 0.7)
    _check_fused(rc, [k], v, filter=rng.random(n) < 0.7, bin_low=0)


def test_fused_groupby_sum_late_keys_nan_and_multikey(rc):
    rng = np.random.default_rng(5)
    n = 6_000_000
    # keys that only appear late: provisional rows inside fused steady-state rounds, summed by the fix-up pass
    k = rng.integers(0, 2000, n).astype(np.int64) * 7919
    late = rng.integers(4_500_000, n, 3000)
    k[late] = 10**12 + rng.integers(0, 500, len(late))
    v = rng.random(n) * 2 - 1
    v[rng.integers(0, n, 1000)] = np.nan
    _check_fused(rc, [k], v, func=50)             # GB_NANSUM skips the NaNs
    u = _check_fused(rc, [k], np.nan_to_num(v))
    assert u > 2000
    # a burst of new keys large enough to overflow the settled table: the fused round aborts, is redone unfused, and
    # the sums are rebuilt from the rows grouped so far
    k2 = k.copy()
    k2[5_000_000:] = 10**15 + np.arange(n - 5_000_000)
    _check_fused(rc, [k2], np.nan_to_num(v))
    # multi-key and string rows have no fused kernel: every round is grouped, then summed
    a = rng.integers(0, 300, 400_000).astype(np.int64)
    b = rng.integers(0, 50, 400_000).astype(np.int32)
    s = np.array([b"k%03d" % i for i in range(200)], dtype="S10")[rng.integers(0, 200, 400_000)]
    w = rng.integers(-1000, 1000, 400_000).astype(np.int32)
    _check_fused(rc, [a, b], w)
    _check_fused(rc, [a, s, b], w.astype(np.float32))
    _check_fused(rc, [s], w.astype(np.float64), filter=rng.random(400_000) < 0.5)


def test_fused_groupby_sum_device_mode_and_empty(rc):
    import torch

    rng = np.random.default_rng(9)
    n = 2_500_000
    k = rng.integers(-(2**62), 2**62, 20_000)[rng.integers(0, 20_000, n)]
    v = rng.random(n)
    ik, ifk, u, s = rc.MultiKeyGroupBySum32([torch.from_numpy(k).cuda()], torch.from_numpy(v).cuda())
    assert ik.is_cuda and ifk.is_cuda and s.is_cuda
    ok, ofk, ou = orc.MultiKeyGroupBy32([k])
    np.testing.assert_array_equal(ik.cpu().numpy(), ok)
    np.testing.assert_array_equal(ifk.cpu().numpy(), ofk)
    exp = orc.GroupByAll32([v], ok, ou, [0], [1], [ou + 1], 0)[0]
    np.testing.assert_allclose(s.cpu().numpy()[1:], exp[1:], rtol=1e-12, atol=1e-10)
    ik, ifk, u, s = rc.MultiKeyGroupBySum32([np.zeros(0, np.int64)], np.zeros(0, np.float32))
    assert u == 0 and len(ik) == 0 and s.dtype == np.float32 and s.tolist() == [0.0]
    with pytest.raises(TypeError):
        rc.MultiKeyGroupBySum32([k], np.array([b"a"] * n))
    with pytest.raises(ValueError):
        rc.MultiKeyGroupBySum32([k], v[:-1])


def test_16_byte_aligned_slices_take_the_direct_path(rc):
    """a 16- but not 32-byte aligned int64 / float64 slice (tensor[2:]) is all the 128-bit loads need; it
    must neither change the slot format of a resumable grouping nor be refused by GroupByAll32 / BinCount."""
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    n = 3_000_000
    base = torch.randint(0, 5000, (n + 2,), dtype=torch.int64, device="cuda", generator=g) * 977 - 12345
    vbase = torch.rand(n + 2, dtype=torch.float64, device="cuda", generator=g)
    k, v = base[2:], vbase[2:]
    assert k.data_ptr() % 32 == 16 and v.data_ptr() % 32 == 16
    kh, vh = k.cpu().numpy(), v.cpu().numpy()
    ok, ofk, ou = orc.MultiKeyGroupBy32([kh])
    ik, ifk, u = rc.MultiKeyGroupBy32([k], 0, None, 2)
    assert u == ou
    np.testing.assert_array_equal(ik.cpu().numpy(), ok)
    np.testing.assert_array_equal(ifk.cpu().numpy(), ofk)
    half = 1_500_000
    st = rc.GroupByStream(1 << 20, 2_000_000)
    a = st.push(k[:half]).clone()
    b = st.push(k[half:], row_offset=half).clone()  # this chunk starts 16 bytes off a 32-byte boundary
    np.testing.assert_array_equal(torch.cat([a, b]).cpu().numpy(), ok)
    np.testing.assert_array_equal(st.ifirstkey.cpu().numpy(), ofk)
    ikl = torch.empty(n + 2, dtype=torch.int64, device="cuda")
    ikl[2:] = ik
    s = rc.GroupByAll32([v], ikl[2:], u, [0], [1], [u + 1], 0)[0]
    exp = orc.GroupByAll32([vh], ok, ou, [0], [1], [ou + 1], 0)[0]
    np.testing.assert_allclose(s.cpu().numpy()[1:], exp[1:], rtol=1e-12, atol=1e-9)
    np.testing.assert_array_equal(rc.BinCount(ikl[2:], u + 1).cpu().numpy(), np.bincount(ok, minlength=ou + 1))
    ik2, ifk2, u2, s2 = rc.MultiKeyGroupBySum32([k], v)
    np.testing.assert_array_equal(ik2.cpu().numpy(), ok)
    np.testing.assert_allclose(s2.cpu().numpy()[1:], exp[1:], rtol=1e-12, atol=1e-9)


def test_groupby_stream_with_sums(rc):
    """The resumable fused call: chunks pushed one after the other == one MultiKeyGroupBySum32 over all rows; a push
    without values (a pre-loaded seed) creates bins whose sums start at zero."""
    import torch

    rng = np.random.default_rng(21)
    n = 5_000_000
    uv = rng.integers(-(2**62), 2**62, 30_000)
    k = uv[rng.integers(0, 30_000, n)]
    # a burst of all-distinct keys that overflows the settled table: the fused round of that chunk aborts, its partial
    # sums are rolled back from the snapshot, and the round is redone unfused
    k[4_000_000:] = 2**62 + 5 * np.arange(1_000_000)
    for vdt in (np.float64, np.float32, np.int32):
        v = (rng.random(n) * 2 - 1).astype(vdt) if np.dtype(vdt).kind == "f" else rng.integers(-50, 50, n).astype(vdt)
        seed = uv[:1000].copy()
        st = rc.GroupByStream(1 << 20, 2_000_000, with_sums=True)
        st.push(torch.from_numpy(seed).cuda())
        assert st.unique_count == 1000
        parts, cuts = [], [0, 1_999_872, 3_999_744, n]
        for a, b in zip(cuts[:-1], cuts[1:]):
            parts.append(st.push(torch.from_numpy(k[a:b]).cuda(), None, a, None, values=torch.from_numpy(v[a:b]).cuda()).clone())
        ok, ofk, ou = orc.MultiKeyGroupBy32([np.concatenate([seed, k])])
        np.testing.assert_array_equal(torch.cat(parts).cpu().numpy(), ok[1000:])
        assert st.unique_count == ou
        exp = orc.GroupByAll32([v], ok[1000:], ou, [0], [1], [ou + 1], 0)[0]
        got = st.sums().cpu().numpy()
        if np.dtype(vdt).kind == "f":
            np.testing.assert_allclose(got[1:], exp[1:].astype(np.float64), rtol=1e-5 if vdt == np.float32 else 1e-12, atol=1e-9 if vdt == np.float64 else 1e-3)
        else:
            np.testing.assert_array_equal(got[1:], exp[1:])


@pytest.mark.parametrize("n", [0, 1, 1000, 300_001])
def test_ismember_categorical_fixup(rc, n):
    rng = np.random.default_rng(n + 5)
    for ba in (0, 1):
        for bb in (0, 1):
            nua, nub = 40, 25
            for adt in (np.int8, np.int32, np.int64):
                a = rng.integers(ba - (1 if ba else 0), nua + ba, n).astype(adt)   # base 1: includes the filtered code 0
                b = rng.integers(0, nub + bb, max(n // 2, 3)).astype(np.int16)
                b[b == 7 + bb] = 3 + bb                                            # one category of b never occurs
                if n > 10:
                    a[5] = np.iinfo(adt).min                                       # a sentinel code
                on_unique = rng.integers(-1, nub, nua).astype(np.int32)
                on_unique[on_unique < 0] = np.iinfo(np.int32).min
                gf, gi = rc.IsMemberCategoricalFixup(a, b, on_unique, nub, ba, bb)
                ef, ei = orc.IsMemberCategoricalFixup(a, b, on_unique, nub, ba, bb)
                assert gf.dtype == np.bool_ and gi.dtype == np.int32
                np.testing.assert_array_equal(gf, ef)
                np.testing.assert_array_equal(gi, ei)
    c1 = np.array([0, -128, 3], dtype=np.int8)
    c2 = np.array([1, 2, -128, 0, 3], dtype=np.int8)
    is_in, idx = rc.IsMemberCategoricalFixup(c1, c2, np.array([0, 1, 2], dtype=np.int32), 4, 1, 1)
    assert is_in.tolist() == [False, False, True] and idx.tolist() == [np.iinfo(np.int32).min] * 2 + [4]


def test_reduce_with_up_to_16k_bins_in_shared_memory(rc):
    """3.5 K - 16 K bins (the reference benchmark grid's 10 000-unique column) keep their accumulators in shared memory
    through the opt-in limit once the column is long enough: every op against the oracle."""
    rng = np.random.default_rng(31)
    n, u = 12_000_000, 10_000
    ik = rng.integers(0, u + 1, n).astype(np.int16)
    v = rng.random(n) * 2 - 1
    v[rng.integers(0, n, 2000)] = np.nan
    vi = rng.integers(-1000, 1000, n).astype(np.int32)
    for fn in (0, 1, 2, 3, 4, 50, 51, 52, 55):
        got = rc.GroupByAll32([v], ik, u, [fn], [1], [u + 1], 0)[0]
        exp = orc.GroupByAll32([v], ik, u, [fn], [1], [u + 1], 0)[0]
        np.testing.assert_allclose(got[1:], exp[1:], rtol=1e-12, atol=1e-9, equal_nan=True, err_msg=f"func {fn}")
    for fn in (0, 2, 3):
        got = rc.GroupByAll32([vi], ik, u, [fn], [0], [u + 1], 0)[0]
        exp = orc.GroupByAll32([vi], ik, u, [fn], [0], [u + 1], 0)[0]
        np.testing.assert_array_equal(got, exp)
    np.testing.assert_array_equal(rc.BinCount(ik, u + 1), np.bincount(ik, minlength=u + 1))


@pytest.mark.parametrize("n", [1, 31, 1000, 65_537, 1_048_576])
def test_small_inputs_single_launch_path(rc, n):
    """Inputs up to 1 Mi rows are grouped by one cooperative launch: same results as the oracle for every key width,
    filters, the key that equals the table's empty marker, all-distinct keys, and a unique-count hint that is too small."""
    rng = np.random.default_rng(n)
    for dt in (np.int64, np.int32, np.uint16, np.int8, np.float64):
        k = rng.integers(0, 100 if np.dtype(dt).itemsize == 1 else max(2, n // 3), n).astype(dt)
        _check_gb(rc, [k])
        _check_gb(rc, [k], filter=rng.random(n) < 0.5)
    k = rng.integers(-3, 3, n).astype(np.int64)  # includes -1, the empty marker
    _check_gb(rc, [k])
    _check_gb(rc, [np.arange(n, dtype=np.int64)[::-1].copy()])          # every row a new key
    _check_gb(rc, [rng.permutation(n).astype(np.int32)], hint=max(1, n // 50))  # hint too small: retried with more room
    a = rng.integers(0, 20, n).astype(np.int16)
    b = rng.integers(0, 20, n).astype(np.int32)
    _check_gb(rc, [a, b])                                                # composed into one 64-bit key
    f, i = rc.IsMember32(rng.integers(0, 50, 1000).astype(np.int64), k)
    ef, ei = orc.IsMember32(rng.integers(0, 50, 0).astype(np.int64), k) if False else (None, None)
    v = rng.random(n)
    _check_fused(rc, [k], v)


@pytest.mark.parametrize("kdt", [np.int64, np.int32, np.uint8])
@pytest.mark.parametrize("with_sums", [False, True])
def test_groupby_stream_adopt_seed(rc, kdt, with_sums):
    """rtb_multikey_groupby32_adopt_seed: a stream that numbered a prefix locally is re-labelled with a seed in place.
    Afterwards it must behave exactly like a stream that had been pre-loaded with the seed: the oracle groups
    [seed ; prefix ; rest] as one array."""
    import torch

    rng = np.random.default_rng(12)
    small = np.dtype(kdt).itemsize == 1
    pool = (np.arange(200) if small else rng.integers(-(2**30), 2**30, 6000)).astype(kdt)
    npre, nrest = 1_500_000, 2_000_000
    prefix = pool[rng.integers(0, len(pool) * 2 // 3, npre)]
    rest = pool[rng.integers(0, len(pool), nrest)]
    # the seed: some keys the prefix has, some it lacks, in an order of its own; the key -1 (empty marker for int64) included
    seed = rng.permutation(np.unique(np.r_[pool[len(pool) // 3:], np.array([-1 if not small else 255]).astype(kdt)]))
    vpre, vrest = rng.random(npre), rng.random(nrest)
    st = rc.GroupByStream(1 << 20, 2_000_000, with_sums=with_sums)
    kw = {"values": torch.from_numpy(vpre).cuda()} if with_sums else {}
    ik_pre = st.push(torch.from_numpy(prefix).cuda(), None, 0, None, **kw).clone()
    u_loc = st.unique_count
    mp = st.adopt_seed(torch.from_numpy(seed).cuda())
    assert mp is not None and mp.numel() == u_loc + 1 and int(mp[0]) == 0
    ik_pre = mp[ik_pre.long()]
    kw = {"values": torch.from_numpy(vrest).cuda()} if with_sums else {}
    ik_rest = st.push(torch.from_numpy(rest).cuda(), None, npre, None, **kw)
    allk = np.concatenate([seed, prefix, rest])
    ok, ofk, ou = orc.MultiKeyGroupBy32([allk])
    u0 = len(seed)
    np.testing.assert_array_equal(ok[:u0], np.arange(1, u0 + 1))
    np.testing.assert_array_equal(ik_pre.cpu().numpy(), ok[u0:u0 + npre])
    np.testing.assert_array_equal(ik_rest.cpu().numpy(), ok[u0 + npre:])
    assert st.unique_count == ou
    # first rows of the keys outside the seed are the shard's own rows
    ifk = st.ifirstkey.cpu().numpy()
    np.testing.assert_array_equal(ifk[u0:], ofk[u0:] - u0)
    if with_sums:
        exp = orc.GroupByAll32([np.concatenate([vpre, vrest])], ok[u0:], ou, [0], [1], [ou + 1], 0)[0]
        np.testing.assert_allclose(st.sums().cpu().numpy()[1:], exp[1:], rtol=1e-12, atol=1e-9)


def test_host_mode_chunk_pipeline(rc, monkeypatch):
    """Large host-mode calls are cut into chunks so that the upload of the next rows, the grouping of the current ones and
    the read-back of the previous iKey overlap (PCIe both ways).  With the thresholds lowered the pipeline runs on a few
    million rows: same results as the oracle, for MultiKeyGroupBy32 and the fused call, with filters and late keys."""
    monkeypatch.setattr(rc, "_PIPE_MIN_ROWS", 100_000)
    monkeypatch.setattr(rc, "_PIPE_CHUNK_ROWS", 256 * 1024)
    rng = np.random.default_rng(44)
    n = 3_300_001
    for kdt, vdt in ((np.int64, np.float64), (np.int32, np.float32), (np.uint16, np.int32), (np.float64, np.int64)):
        k = rng.integers(0, 40_000, n).astype(kdt)
        k[n - 400_000:] = (rng.integers(0, 60_000, 400_000)).astype(kdt)  # keys that first appear in the last chunks
        v = (rng.random(n) * 2 - 1).astype(vdt) if np.dtype(vdt).kind == "f" else rng.integers(-100, 100, n).astype(vdt)
        f = rng.random(n) < 0.8
        for filt in (None, f):
            ok, ofk, ou = orc.MultiKeyGroupBy32([k], 0, filt, 2)
            ik, ifk, u = rc.MultiKeyGroupBy32([k], 0, filt, 2)
            assert u == ou and ik.dtype == np.int32 and ifk.dtype == np.int32
            np.testing.assert_array_equal(ik, ok)
            np.testing.assert_array_equal(ifk, ofk)
            _check_fused(rc, [k], v, filter=filt)
    # a unique-count hint that is too small: the pipeline gives up and the one-shot path retries with more room
    k = rng.permutation(n).astype(np.int64)
    ik, ifk, u = rc.MultiKeyGroupBy32([k], 1000, None, 2)
    assert u == n and np.array_equal(ifk, np.arange(n))
    # pinned inputs take the direct copies
    import torch

    kp = torch.from_numpy(rng.integers(0, 5000, n)).pin_memory()
    vp = torch.rand(n, dtype=torch.float64).pin_memory()
    _check_fused(rc, [kp.numpy()], vp.numpy())


@pytest.mark.parametrize("cols", ["i64_s16_i32", "s12", "i64_i64", "s40_i32"])
def test_packed_rows_deferred_general_path(rc, monkeypatch, cols):
    """RTB_GB_DEFER_BYTES=0: packed-row rounds after the first run as two launches whatever the table size -- the fast path
    (home bucket, representative row, compare) and a second launch over the rows it marked (collisions, provisional slots,
    new keys).  Many uniques, a filter, rows that differ only in the last column."""
    monkeypatch.setenv("RTB_GB_DEFER_BYTES", "0")
    rng = np.random.default_rng(41)
    n = 2_500_003
    a = rng.integers(0, 3000, n).astype(np.int64) * 1_000_003
    b = rng.integers(0, 50, n).astype(np.int32)
    pool = np.array([("%c%c" % (65 + i % 26, 65 + i // 26)).encode() * 6 for i in range(300)], dtype="S12")
    s12 = pool[rng.integers(0, 300, n)]
    if cols == "i64_s16_i32":
        keys = [a, s12.astype("S16"), b]
    elif cols == "s12":
        keys = [pool[rng.integers(0, 300, n)].astype("S12")]
        keys = [np.char.add(keys[0], rng.integers(0, 2000, n).astype("S4")).astype("S12")]
    elif cols == "i64_i64":
        keys = [a, rng.integers(0, 40, n).astype(np.int64)]
    else:
        keys = [np.char.add(s12, rng.integers(0, 500, n).astype("S6")).astype("S40"), b]
    filt = rng.random(n) < 0.9
    for f in (None, filt):
        ik, ifk, u = rc.MultiKeyGroupBy32(keys, 0, f, 2)
        eik, eifk, eu = orc.MultiKeyGroupBy32(keys, 0, f, 2)
        assert u == eu
        np.testing.assert_array_equal(ik, eik)
        np.testing.assert_array_equal(ifk, eifk)


@pytest.mark.parametrize("uniques", [2_200_000, 3_500_000])
def test_table_window_sizing(rc, uniques):
    """Unique counts whose 25 %-load table would leave the 64 MB window: the table stays at 4 Mi slots (2.2 M keys, 52 % load)
    or is sized for 50 % load beyond it (3.5 M keys, 8 Mi slots) -- same bins, first rows and count as the oracle."""
    rng = np.random.default_rng(uniques)
    n = 14_000_000
    key = rng.integers(0, uniques, n).astype(np.int64) * 104_729
    ik, ifk, u = rc.MultiKeyGroupBy32([key], 0, None, 2)
    eik, eifk, eu = orc.MultiKeyGroupBy32([key], 0, None, 2)
    assert u == eu
    np.testing.assert_array_equal(ik, eik)
    np.testing.assert_array_equal(ifk, eifk)
