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


def test_c_parallel_groupby_equals_sequential():
    # the cutoffs + hstack_groupings route (partitions grouped in parallel, merged in partition order)
    rng = np.random.default_rng(3)
    vals = rng.integers(-(2**62), 2**62, 5000)
    k = vals[rng.integers(0, 5000, 400_000)]
    k[:200_000] = vals[rng.integers(0, 2000, 200_000)]  # late keys
    f = rng.random(len(k)) < 0.8
    for filt in (None, f):
        ok, ofk, ou = orc.MultiKeyGroupBy32([k], filter=filt)
        for threads in (2, 5, 8):
            ik, ifk, u = c_oracle.groupby_u64(k, filt, threads=threads)
            assert u == ou and np.array_equal(ik, ok) and np.array_equal(ifk, ofk)
