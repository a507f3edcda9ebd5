"""A compact tour of the hot-path kernels for compute-sanitizer (memcheck / racecheck): every entry point once or twice at
sizes that reach the tiled steady-state kernels, results checked against the oracle so a silent wrong answer also fails."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import riptable_b200.rc as rc
from riptable_b200 import merge
from oracle import riptide_oracle as orc
from oracle import merge_oracle as mo

rng = np.random.default_rng(1)
n = int(os.environ.get("SAN_ROWS", 2_300_000))  # > 1 Mi rows: the round schedule and gb_round_direct run
k = rng.integers(-(2**40), 2**40, 3000)[rng.integers(0, 3000, n)]
k[n - 50_000:] = 2**50 + rng.integers(0, 200, 50_000)  # late keys
v = rng.random(n) * 2 - 1
ok, ofk, ou = orc.MultiKeyGroupBy32([k])
ik, ifk, u = rc.MultiKeyGroupBy32([k], 0, None, 2)
assert u == ou and np.array_equal(ik, ok) and np.array_equal(ifk, ofk)
ik2, ifk2, u2, s2 = rc.MultiKeyGroupBySum32([k], v)
assert np.array_equal(ik2, ok)
exp = orc.GroupByAll32([v], ok, ou, [0], [1], [ou + 1], 0)[0]
np.testing.assert_allclose(s2[1:], exp[1:], rtol=1e-12, atol=1e-9)
m = 200_000
ks, vs = k[:m], v[:m]
oks, ofs, ous = orc.MultiKeyGroupBy32([ks], 0, rng.random(m) < 0.7, 2)
f = rng.random(m) < 0.7
oks, ofs, ous = orc.MultiKeyGroupBy32([ks], 0, f, 2)
g = rc.MultiKeyGroupBy32([ks], 0, f, 2)                       # small single-launch path
assert np.array_equal(g[0], oks) and np.array_equal(g[1], ofs)
a32 = rng.integers(0, 300, m).astype(np.int32)
s10 = np.array([b"k%04d" % i for i in range(500)], dtype="S10")[rng.integers(0, 500, m)]
for keys in ([a32, s10], [s10], [ks, a32]):
    e = orc.MultiKeyGroupBy32(keys)
    g = rc.MultiKeyGroupBy32(keys, 0, None, 2)
    assert np.array_equal(g[0], e[0]) and np.array_equal(g[1], e[1])
cut = np.array([50_000, 120_000, m], dtype=np.int64)
e = orc.MultiKeyGroupBy32([ks], cutoffs=cut)
g = rc.MultiKeyGroupBy32([ks], 0, None, 2, cutoffs=cut)
assert np.array_equal(g[0], e[0]) and np.array_equal(g[1][0], e[1][0])
iks = oks.astype(np.int16) if ous < 30000 else oks
for fn in (0, 1, 2, 3, 4, 51, 53):
    got = rc.GroupByAll32([vs], iks, ous, [fn], [1], [ous + 1], 0)[0]
    ex = orc.GroupByAll32([vs], iks, ous, [fn], [1], [ous + 1], 0)[0]
    np.testing.assert_allclose(got[1:], ex[1:], rtol=1e-12, atol=1e-9, equal_nan=True)
cnt, ig, ifg = rc.BinCount(oks, ous + 1, pack=True)
ec, eg, ef = orc.BinCount(oks, ous + 1, pack=True)
assert np.array_equal(ig, eg) and np.array_equal(ifg, ef) and np.array_equal(cnt, ec)
for fn in (100, 102):
    assert np.array_equal(rc.GroupByAllPack32([vs], oks, ig, ifg, cnt, ous, [fn], [1], [ous + 1], 0)[0][1:],
                          orc.GroupByAllPack32([vs], oks, eg, ef, ec, ous, [fn], [1], [ous + 1], 0)[0][1:])
np.testing.assert_allclose(rc.EmaAll32([vs], oks, ous, 300, (0.0, None, None, None))[0],
                           orc.EmaAll32([vs], oks, ous, 300, (0.0, None, None, None))[0], rtol=1e-12, atol=1e-9, equal_nan=True)
t = np.cumsum(rng.integers(0, 3, 20_000)).astype(np.float64)
np.testing.assert_allclose(rc.EmaAll32([vs[:20_000]], oks[:20_000], ous, 304, (0.2, t, None, None))[0],
                           orc.EmaAll32([vs[:20_000]], oks[:20_000], ous, 304, (0.2, t, None, None))[0], rtol=1e-10, atol=1e-9, equal_nan=True)
fl = np.round(rng.random(m) * 30)
fl[rng.random(m) < 0.02] = np.nan
p = rc.LexSort32([a32, fl, s10])
assert np.array_equal(p, np.lexsort([a32, fl, s10]))
assert np.array_equal(rc.LexSort32([a32, fl], cutoffs=cut), orc.LexSort32([a32, fl], cutoffs=cut))
g = rc.GroupFromLexSort(p, [a32, fl, s10])
e = orc.GroupFromLexSort(p, [a32, fl, s10])
assert all(np.array_equal(x, y) for x, y in zip(g, e))
fnd, idx = rc.IsMember32(ks, k[:5000])
efnd, eidx = orc.IsMember32(ks, k[:5000])
assert np.array_equal(fnd, efnd) and np.array_equal(idx[fnd], eidx[efnd])
for how in ("left", "inner", "outer"):
    gl, gr, gk = merge.merge_indices([a32[:30_000]], [a32[30_000:50_000] + 100], how)
    el, er, ek = mo.merge_indices([a32[:30_000]], [a32[30_000:50_000] + 100], how)
    assert gk == ek and (gr is None) == (er is None)
    if er is not None:
        assert np.array_equal(np.where(gr == merge.INV, -1, gr), np.where(er == mo.INV, -1, er))
assert np.array_equal(rc.MakeiNext(oks, ous, 0), orc.MakeiNext(oks, ous, 0))
print("sanitize tour ok")
