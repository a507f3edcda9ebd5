"""GPU: BASELINE.json's full-size configurations checked through size-independent properties (the oracle cannot run
at these sizes in seconds): first-appearance invariants, conservation laws, sortedness, idempotence, and agreement
between independent paths (hash vs lexsort grouping, pack vs bincount, cumsum's last value vs sum)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rc():
    import riptable_b200.rc as rc

    return rc


def _c2(torch, n, u, seed=12345):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    uvals = torch.randint(-(2**63), 2**63 - 1, (u,), dtype=torch.int64, device="cuda", generator=g)
    key = torch.empty(n, dtype=torch.int64, device="cuda")
    chunk = 1 << 27
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        idx = torch.randint(0, u, (e - s,), dtype=torch.int64, device="cuda", generator=g)
        torch.index_select(uvals, 0, idx, out=key[s:e])
    return key, uvals


def test_c2_groupby_sum_min_max_count_1e9_rows(rc):
    """configs[1]: 1 B rows, int64 key, 1 M uniques, f32 + f64 values."""
    import torch

    n, u = 1_000_000_000, 1_000_000
    key, uvals = _c2(torch, n, u)
    ik, ifk, U = rc.MultiKeyGroupBy32([key], 0, None, 2)
    assert U == int(torch.unique(uvals).numel())
    assert ik.dtype == torch.int32 and int(ik.min()) == 1 and int(ik.max()) == U
    # iFirstKey strictly increasing; the row it names carries bin b and no earlier row does (first appearance)
    assert bool((ifk[1:] > ifk[:-1]).all())
    assert bool((ik[ifk.long()] == torch.arange(1, U + 1, device="cuda", dtype=torch.int32)).all())
    first_seen = torch.full((U + 1,), n, dtype=torch.int64, device="cuda")
    first_seen.scatter_reduce_(0, ik.long(), torch.arange(n, device="cuda"), reduce="amin")
    assert bool((first_seen[1:] == ifk.long()).all())
    # same key <=> same bin: the key of every row equals the key of its bin's first row
    step = 7
    rows = torch.arange(0, n, step, device="cuda")
    assert bool((key[rows] == key[ifk.long()[ik[rows].long() - 1]]).all())
    cnt = rc.BinCount(ik, U + 1)
    assert int(cnt.sum()) == n and int(cnt[0]) == 0
    for dt, rtol in ((torch.float64, 1e-12), (torch.float32, 1e-5)):
        g = torch.Generator(device="cuda")
        g.manual_seed(7)
        val = torch.empty(n, dtype=dt, device="cuda").uniform_(-1.0, 1.0, generator=g)
        s = rc.GroupByAll32([val], ik, U, [0], [1], [U + 1], 0)[0]
        assert s.dtype == dt
        ref = torch.zeros(U + 1, dtype=torch.float64, device="cuda").index_add_(0, ik.long(), val.double())
        scale = float(cnt.max()) ** 0.5
        assert float((s.double() - ref)[1:].abs().max()) <= rtol * max(scale, 1.0) * 8
        mn = rc.GroupByAll32([val], ik, U, [2], [1], [U + 1], 0)[0]
        mx = rc.GroupByAll32([val], ik, U, [3], [1], [U + 1], 0)[0]
        rmn = torch.full((U + 1,), float("inf"), dtype=dt, device="cuda").scatter_reduce_(0, ik.long(), val, reduce="amin")
        rmx = torch.full((U + 1,), float("-inf"), dtype=dt, device="cuda").scatter_reduce_(0, ik.long(), val, reduce="amax")
        assert bool((mn[1:] == rmn[1:]).all()) and bool((mx[1:] == rmx[1:]).all())  # bit-exact
        del val, ref, rmn, rmx
    # idempotence: grouping the bins reproduces them
    ik2, ifk2, U2 = rc.MultiKeyGroupBy32([ik], 0, None, 2)
    assert U2 == U and bool((ik2 == ik).all()) and bool((ifk2 == ifk).all())


def _categorical_route(rcm, key, val, enums):
    """`Categorical(key).sum(val) / .mean(val)` restated (riptable/rt_grouping.py:944-987, 1009-1043): hash grouping,
    uniques sorted, iKey re-indexed through ismember to the sorted order, down-cast by unique count, then the
    per-bin reductions against that iKey."""
    ik, ifk, u = rcm.MultiKeyGroupBy32([key], 0, None, 2)
    uniques = key[ifk]
    sortidx = rcm.LexSort32([uniques])                                   # _make_isortrows (rt_grouping.py:2204-2210)
    _, ik = rcm.IsMemberCategorical(ik, (sortidx + 1).astype(ik.dtype))  # rt_grouping.py:963
    ik = ik.astype(enums.int_dtype_from_len(u))                          # possibly_recast (rt_enum.py:666-679)
    s = rcm.GroupByAll32([val], ik, u, [0], [1], [u + 1], 0)[0]
    m = rcm.GroupByAll32([val], ik, u, [1], [1], [u + 1], 0)[0]
    return ik, uniques[sortidx], u, s, m


def test_c1_categorical_sum_mean_10m_rows_against_the_oracle(rc):
    """configs[0] at its stated size, row for row against the oracle: 10 M rows, int32 key with 1 K uniques, f64 value
    (recipe of SURVEY 8d: seed 12345, value = random * 1e10 - 0.5e10), through the Categorical route: the stored iKey
    is int16 and the reductions run on the shared-memory path."""
    from oracle import riptide_oracle as orc
    from riptable_b200 import enums

    rng = np.random.default_rng(12345)
    n, u = 10_000_000, 1000
    key = np.arange(u, dtype=np.int32)[rng.integers(0, u, n)]
    val = rng.random(n) * 1e10 - 0.5e10
    gik, gcat, gu, gs, gm = _categorical_route(rc, key, val, enums)
    oik, ocat, ou, os_, om = _categorical_route(orc, key, val, enums)
    assert gu == ou == u and gik.dtype == np.int16 == oik.dtype
    np.testing.assert_array_equal(gik, oik)                                   # every row's bin, bit-exact
    np.testing.assert_array_equal(gcat, ocat)
    np.testing.assert_array_equal(gik, (np.searchsorted(gcat, key) + 1))      # ordered Categorical: bin = rank of the key
    assert gs.dtype == np.float64 and gm.dtype == np.float64 and gs.shape == (u + 1,)
    # 1e-12 relative to the data magnitude (the values are ~1e10 and mostly cancel inside a bin)
    scale = np.abs(val).max() * np.sqrt(n / u)
    np.testing.assert_allclose(gs[1:], os_[1:], rtol=1e-12, atol=1e-12 * scale)
    np.testing.assert_allclose(gm[1:], om[1:], rtol=1e-12, atol=1e-12 * scale / (n / u))
    np.testing.assert_allclose(gs[1:], np.bincount(gik, weights=val, minlength=u + 1)[1:], rtol=1e-12, atol=1e-12 * scale)
    # the fused entry on the same column gives the unordered (first-appearance) grouping and the same sums, re-ordered
    fik, fifk, fu, fs = rc.MultiKeyGroupBySum32([key], val)
    eik, eifk, eu = orc.MultiKeyGroupBy32([key])
    np.testing.assert_array_equal(fik, eik)
    np.testing.assert_array_equal(fifk, eifk)
    order = np.searchsorted(gcat, key[fifk]) + 1                               # first-appearance bin -> sorted bin
    np.testing.assert_allclose(fs[1:], os_[order], rtol=1e-12, atol=1e-12 * scale)


def test_c3_multikey_with_strings_nanmean_nanstd(rc):
    """configs[2] at its stated size: 500 M rows, keys (int64 of 1e4 values, S16 of 1e3 strings, int32 of 5 values)
    -> about 5e7 distinct tuples, value f64 with 1 % NaN, nanmean / nanstd."""
    import torch

    n = 500_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    k1 = torch.randint(0, 10_000, (n,), device="cuda", generator=g, dtype=torch.int64) * 1_000_003
    pool = np.frombuffer(np.random.default_rng(1).integers(65, 91, 1000 * 16, dtype=np.uint8).tobytes(), dtype="S16")
    sidx = torch.randint(0, 1000, (n,), device="cuda", generator=g, dtype=torch.int64)
    s16 = rc.DevArray(torch.from_numpy(pool.view(np.uint8).reshape(1000, 16).copy()).cuda()[sidx].reshape(-1).contiguous(),
                      np.dtype("S16"), n)
    k3 = torch.randint(0, 5, (n,), device="cuda", generator=g, dtype=torch.int32)
    ik, ifk, U = rc.MultiKeyGroupBy32([k1, s16, k3], 0, None, 2)
    # independent count of distinct tuples: combine the three ids into one int64 and count uniques
    combo = (k1 // 1_000_003) * 5000 + sidx * 5 + k3.long()
    assert U == int(torch.unique(combo).numel())
    assert bool((ifk[1:] > ifk[:-1]).all())
    assert bool((combo == combo[ifk.long()[ik.long() - 1]]).all())
    assert 45_000_000 <= U <= 50_000_000
    del combo, k1, s16, k3, sidx
    torch.cuda.empty_cache()
    val = torch.rand(n, device="cuda", generator=g, dtype=torch.float64)
    val[torch.rand(n, device="cuda", generator=g) < 0.01] = float("nan")
    mean = rc.GroupByAll32([val], ik, U, [51], [1], [U + 1], 0)[0]
    std = rc.GroupByAll32([val], ik, U, [55], [1], [U + 1], 0)[0]
    ok = ~torch.isnan(val)
    c = torch.zeros(U + 1, dtype=torch.float64, device="cuda").index_add_(0, ik.long()[ok], torch.ones(int(ok.sum()), dtype=torch.float64, device="cuda"))
    sm = torch.zeros(U + 1, dtype=torch.float64, device="cuda").index_add_(0, ik.long()[ok], val[ok])
    rmean = sm / c
    d = val[ok] - rmean[ik.long()[ok]]
    m2 = torch.zeros(U + 1, dtype=torch.float64, device="cuda").index_add_(0, ik.long()[ok], d * d)
    rstd = torch.where(c >= 2, torch.sqrt(m2 / (c - 1)), torch.full_like(m2, float('nan')))  # ddof = 1; < 2 valid rows -> NaN
    torch.testing.assert_close(mean[1:], rmean[1:], rtol=1e-12, atol=1e-12, equal_nan=True)
    torch.testing.assert_close(std[1:], rstd[1:], rtol=1e-10, atol=1e-12, equal_nan=True)


def test_c4_lexsort_three_keys_and_group_from_lexsort(rc):
    """configs[3] at its stated size: 1 B rows, sort keys (int64 of 1e3 values, f64 with 0.1 % NaN, int32 of 1e6
    values), int64 primary; then GroupFromLexSort.  Verified in 100 M-row slices of the sorted order (sortedness under
    np.lexsort's rules incl. NaN last, stability, permutation, bin boundaries) to keep the checker's memory bounded."""
    import torch

    n = 1_000_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    a = torch.randint(0, 1000, (n,), device="cuda", generator=g, dtype=torch.int64) - 500
    f = torch.rand(n, device="cuda", generator=g, dtype=torch.float64) - 0.5
    f[torch.rand(n, device="cuda", generator=g) < 0.001] = float("nan")
    c = torch.randint(0, 1_000_000, (n,), device="cuda", generator=g, dtype=torch.int32)
    perm = rc.LexSort32([c, f, a])  # last key (a) primary
    assert perm.dtype == torch.int32 and perm.numel() == n
    chk = torch.zeros(n, dtype=torch.bool, device="cuda")
    chk[perm.long()] = True
    assert bool(chk.all())  # a permutation
    del chk
    ik, ifk, cnt = rc.GroupFromLexSort(perm, [a, f, c], base_index=1)
    U = int(ifk.numel())
    assert int(cnt.sum()) == n and int(cnt[0]) == 0 and cnt.numel() == U + 1
    step = 100_000_000
    nbins = 0
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        p = perm[max(lo - 1, 0):hi].long()  # one row of overlap so that every adjacent pair is checked
        sa, sf, sc = a[p], f[p], c[p]
        nan2 = torch.isnan(sf)
        fkey = torch.where(nan2, torch.full_like(sf, float("inf")), sf)
        same_a = sa[1:] == sa[:-1]
        f_le = (fkey[1:] >= fkey[:-1]) & ~(nan2[:-1] & ~nan2[1:])
        same_f = (sf[1:] == sf[:-1]) | (nan2[1:] & nan2[:-1])
        same_c = sc[1:] == sc[:-1]
        assert bool((sa[1:] >= sa[:-1]).all())
        assert bool((~same_a | f_le).all())
        assert bool((~(same_a & same_f) | (sc[1:] >= sc[:-1])).all())
        assert bool((~(same_a & same_f & same_c) | (p[1:] > p[:-1])).all())  # ties keep row order
        new_group = ~(same_a & same_f & same_c)
        if lo == 0:
            new_group = torch.cat([torch.ones(1, dtype=torch.bool, device="cuda"), new_group])
        bins = nbins + torch.cumsum(new_group.int(), 0)
        assert bool((ik[perm[lo:hi].long()] == bins.int()).all())  # bins numbered in sorted order, base 1
        nbins = int(bins[-1])
        del p, sa, sf, sc, nan2, fkey, same_a, f_le, same_f, same_c, new_group, bins
    assert U == nbins
    assert bool((ik[ifk.long()] == torch.arange(1, U + 1, device="cuda", dtype=torch.int32)).all())


def test_c5_high_cardinality_pack_first_last_cumsum(rc):
    """configs[4] at its stated size: 1 B rows, 100 M uniques: hash grouping, pack, first / last, cumsum."""
    import torch

    n, u = 1_000_000_000, 100_000_000
    key, uvals = _c2(torch, n, u, seed=5)
    ik, ifk, U = rc.MultiKeyGroupBy32([key], 0, None, 2)
    assert U == int(torch.unique(uvals).numel()) or U <= u  # (nearly) every one of the 1e8 values occurs in 1e9 draws
    assert bool((ifk[1:] > ifk[:-1]).all())
    assert bool((ik[ifk.long()] == torch.arange(1, U + 1, device="cuda", dtype=torch.int32)).all())
    rows = torch.arange(0, n, 5, device="cuda")
    assert bool((key[rows] == key[ifk.long()[ik[rows].long() - 1]]).all())  # same key <=> same bin
    del key, rows, uvals
    torch.cuda.empty_cache()
    cnt, ig, ifg = rc.BinCount(ik, U + 1, pack=True)
    assert int(cnt.sum()) == n
    assert bool((ifg[1:] == torch.cumsum(cnt.long(), 0)[:-1].int()).all())
    packed_bins = ik[ig.long()]
    assert bool((packed_bins[1:] >= packed_bins[:-1]).all())                         # grouped by bin
    assert bool(((packed_bins[1:] != packed_bins[:-1]) | (ig[1:] > ig[:-1])).all())  # rows ascending inside a bin
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    val = torch.randint(-1000, 1000, (n,), device="cuda", generator=g, dtype=torch.int64)
    first = rc.GroupByAllPack32([val], ik, ig, ifg, cnt, U, [100], [1], [U + 1], 0)[0]
    last = rc.GroupByAllPack32([val], ik, ig, ifg, cnt, U, [102], [1], [U + 1], 0)[0]
    assert bool((first[1:] == val[ifk.long()]).all())
    lastrow = torch.zeros(U + 1, dtype=torch.int64, device="cuda").scatter_reduce_(0, ik.long(), torch.arange(n, device="cuda"), reduce="amax")
    assert bool((last[1:] == val[lastrow[1:]]).all())
    cs = rc.EmaAll32([val], ik, U, 300, (0.0, None, None, None))[0]
    sums = rc.GroupByAll32([val], ik, U, [0], [1], [U + 1], 0)[0]
    assert bool((cs[lastrow[1:]] == sums[1:]).all())       # running sum at a bin's last row == the bin's sum (exact, int64)
    assert bool((cs[ifk.long()] == val[ifk.long()]).all())  # and at its first row == that row's value
    nxt = rc.MakeiNext(ik, U, 0)
    has = nxt >= 0
    r = torch.arange(n, device="cuda")[has]
    assert bool((ik[nxt[has].long()] == ik[r]).all()) and bool((nxt[has].long() > r).all())
    assert bool((cs[nxt[has].long()] == cs[r] + val[nxt[has].long()]).all())


def test_c2_recipe_1e8_rows_row_for_row_against_the_c_oracle(rc):
    """configs[1]'s recipe at the largest size the CPU restatement finishes in seconds (1e8 rows, 1 M uniques from the full
    int64 range): iKey / iFirstKey bit for bit, sums / min / max / count against the oracle at the 1e-12 relative bar of
    north_star — through the two reference calls and through the fused call (host mode: the chunk pipeline)."""
    from oracle import c_oracle
    from oracle import riptide_oracle as orc

    rng = np.random.default_rng(12345)
    n, u = 100_000_000, 1_000_000
    uvals = rng.integers(-(2**63), 2**63 - 1, u, dtype=np.int64)
    key = uvals[rng.integers(0, u, n)]
    val = rng.random(n) * 2.0 - 1.0
    threads = max(1, len(__import__("os").sched_getaffinity(0)))
    oik, oifk, ou = c_oracle.groupby_u64(key, threads=threads)
    osum = c_oracle.sum_f64(val, oik, ou, threads=threads)
    ik, ifk, U = rc.MultiKeyGroupBy32([key], 0, None, 2)
    assert U == ou
    np.testing.assert_array_equal(ik, oik)
    np.testing.assert_array_equal(ifk, oifk)
    s = rc.GroupByAll32([val], ik, U, [0], [1], [U + 1], 0)[0]
    rel = np.abs(s[1:] - osum[1:]) / np.maximum(1.0, np.abs(osum[1:]))
    assert rel.max() <= 1e-12, rel.max()
    fik, fifk, fu, fs = rc.MultiKeyGroupBySum32([key], val)
    assert fu == ou
    np.testing.assert_array_equal(fik, oik)
    np.testing.assert_array_equal(fifk, oifk)
    rel = np.abs(fs[1:] - osum[1:]) / np.maximum(1.0, np.abs(osum[1:]))
    assert rel.max() <= 1e-12, rel.max()
    # min / max are exact, count is exact; float32 sums at 1e-5
    cnt = rc.BinCount(ik, U + 1)
    np.testing.assert_array_equal(cnt, np.bincount(oik, minlength=ou + 1))
    sub = slice(0, 20_000_000)  # the NumPy oracle's min / max on a 2e7-row prefix of the same grouping
    for fn in (2, 3):
        np.testing.assert_array_equal(rc.GroupByAll32([val[sub]], ik[sub], U, [fn], [1], [U + 1], 0)[0][1:],
                                      orc.GroupByAll32([val[sub]], oik[sub], ou, [fn], [1], [ou + 1], 0)[0][1:])
    v32 = val.astype(np.float32)
    s32 = rc.GroupByAll32([v32], ik, U, [0], [1], [U + 1], 0)[0]
    assert s32.dtype == np.float32
    ref32 = np.bincount(oik, weights=v32.astype(np.float64), minlength=ou + 1)
    assert (np.abs(s32[1:] - ref32[1:]) / np.maximum(1.0, np.abs(ref32[1:]))).max() <= 1e-5
