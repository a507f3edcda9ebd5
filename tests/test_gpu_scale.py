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


def test_c3_multikey_with_strings_nanmean_nanstd(rc):
    """configs[2] at 1/5 scale (100 M rows): keys (int64, S16, int32), value f64 with 1 % NaN."""
    import torch

    n = 100_000_000
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
    """configs[3] at 1/5 scale (200 M rows): sort keys (int64 1e3 values, f64 with NaN, int32 1e6 values)."""
    import torch

    n = 200_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    a = torch.randint(0, 1000, (n,), device="cuda", generator=g, dtype=torch.int64) - 500
    f = torch.rand(n, device="cuda", generator=g, dtype=torch.float64) - 0.5
    f[torch.rand(n, device="cuda", generator=g) < 0.001] = float("nan")
    c = torch.randint(0, 1_000_000, (n,), device="cuda", generator=g, dtype=torch.int32)
    perm = rc.LexSort32([c, f, a])  # last key (a) primary
    p = perm.long()
    assert int(torch.bincount(p[:: max(1, n // 50_000_000)].clamp(0, 0)).sum()) > 0  # touches the result
    chk = torch.zeros(n, dtype=torch.bool, device="cuda")
    chk[p] = True
    assert bool(chk.all())  # a permutation
    del chk
    sa, sf, sc = a[p], f[p], c[p]
    nan2 = torch.isnan(sf)
    fkey = torch.where(nan2, torch.full_like(sf, float("inf")), sf)
    a_ok = sa[1:] >= sa[:-1]
    same_a = sa[1:] == sa[:-1]
    f_le = (fkey[1:] >= fkey[:-1]) & ~(nan2[:-1] & ~nan2[1:])
    same_f = (sf[1:] == sf[:-1]) | (nan2[1:] & nan2[:-1])
    c_ok = sc[1:] >= sc[:-1]
    same_c = sc[1:] == sc[:-1]
    stable = p[1:] > p[:-1]
    assert bool(a_ok.all())
    assert bool((~same_a | f_le).all())
    assert bool((~(same_a & same_f) | c_ok).all())
    assert bool((~(same_a & same_f & same_c) | stable).all())  # ties keep row order
    ik, ifk, cnt = rc.GroupFromLexSort(perm, [a, f, c], base_index=1)
    U = int(ifk.numel())
    assert int(cnt.sum()) == n and int(cnt[0]) == 0 and cnt.numel() == U + 1
    new_group = torch.ones(n, dtype=torch.bool, device="cuda")
    new_group[1:] = ~(same_a & same_f & same_c)
    assert U == int(new_group.sum())
    assert bool((ik[p] == torch.cumsum(new_group.int(), 0)).all())


def test_c5_high_cardinality_pack_first_last_cumsum(rc):
    """configs[4] at 1/5 scale (200 M rows, 20 M uniques): pack, first/last, cumsum."""
    import torch

    n, u = 200_000_000, 20_000_000
    key, uvals = _c2(torch, n, u, seed=5)
    ik, ifk, U = rc.MultiKeyGroupBy32([key], 0, None, 2)
    del key
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
