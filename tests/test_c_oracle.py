"""The C restatement (oracle/oracle.c) agrees with the NumPy oracle, which is pinned to the golden vectors."""
import numpy as np

from oracle import c_oracle, riptide_oracle as orc


def test_c_groupby_matches_numpy_oracle():
    np.random.seed(12345)
    c = np.random.randint(0, 8000, 2_000_000).astype(np.int64)  # rt_numpy.py:2013-2021
    ik, ifk, u = c_oracle.groupby_u64(c)
    assert u == 8000 and ik[-3:].tolist() == [6061, 7889, 3002] and ifk[-3:].tolist() == [67072, 67697, 68250]
    ok, ofk, ou = orc.MultiKeyGroupBy32([c])
    assert np.array_equal(ik, ok) and np.array_equal(ifk, ofk) and u == ou
    f = (c % 3).astype(bool)
    ik, ifk, u = c_oracle.groupby_u64(c, f)
    ok, ofk, ou = orc.MultiKeyGroupBy32([c], filter=f)
    assert u == 5333 and np.array_equal(ik, ok) and np.array_equal(ifk, ofk)
    rng = np.random.default_rng(1)
    k = rng.integers(-(2**62), 2**62, 300_000)
    ik, ifk, u = c_oracle.groupby_u64(k)
    ok, ofk, ou = orc.MultiKeyGroupBy32([k])
    assert np.array_equal(ik, ok) and np.array_equal(ifk, ofk)


def test_c_sum_matches_numpy_oracle():
    rng = np.random.default_rng(2)
    n, u = 500_000, 1000
    ikey = rng.integers(0, u + 1, n).astype(np.int32)
    v = rng.random(n) * 2 - 1
    for threads in (1, 3, 8):
        got = c_oracle.sum_f64(v, ikey, u, threads=threads)
        exp = orc.GroupByAll32([v], ikey, u, [0], [1], [u + 1])[0]
        np.testing.assert_allclose(got[1:], exp[1:], rtol=1e-12, atol=1e-12)
        assert got[0] == 0
