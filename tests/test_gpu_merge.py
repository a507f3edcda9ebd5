"""GPU parity for SURVEY 8f.4: riptable_b200.merge.merge_indices (the device mirror of rt.merge_indices,
riptable/rt_merge.py:2080-2292) against oracle/merge_oracle.py, which restates the reference's join-index construction and
is itself pinned to pandas.merge in tests/test_oracle_golden.py.  Fancy indices are compared value for value (the invalid
of either index dtype counts as "no row")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import merge_oracle as mo


def _norm(idx, inv):
    if idx is None:
        return None
    out = np.asarray(idx).astype(np.int64)
    out[np.asarray(idx) == inv] = -1
    return out


def _check(lk, rk, how, keep):
    from riptable_b200 import merge

    gl, gr, gk = merge.merge_indices(lk, rk, how, keep)
    el, er, ek = mo.merge_indices(lk, rk, how, keep)
    assert (gl is None) == (el is None) and (gr is None) == (er is None), (how, keep)
    for g, e in ((gl, el), (gr, er)):
        if e is not None:
            assert g.dtype == np.int32
            np.testing.assert_array_equal(_norm(g, merge.INV), _norm(e, mo.INV), err_msg=f"{how} {keep}")
    assert gk == ek
    return gl, gr


KEEPS = [(None, None), (None, "first"), (None, "last"), ("first", None), ("last", None), ("first", "last"), ("last", "first")]


@pytest.mark.parametrize("how", ["left", "right", "inner", "outer"])
@pytest.mark.parametrize("keep", KEEPS)
def test_merge_indices_single_key(how, keep):
    if how == "outer" and keep[0] is not None:
        with pytest.raises(ValueError):
            from riptable_b200 import merge
            merge.merge_indices([np.arange(3)], [np.arange(3)], how, keep)
        return
    rng = np.random.default_rng(abs(hash((how, keep))) % 2**32)
    for nl, nr, span in ((0, 5, 3), (7, 0, 3), (1, 1, 1), (50, 40, 12), (3000, 5000, 400), (2000, 300, 5000), (40_000, 30_000, 50)):
        lk, rk = rng.integers(0, span, nl), rng.integers(span // 3, span + span // 3 + 1, nr)
        _check([lk], [rk], how, keep)
    # float keys with NaN: NaN never matches (SQL JOIN semantics) but the rows stay where the join keeps unmatched rows
    lk = np.round(rng.random(5000) * 50)
    rk = np.round(rng.random(4000) * 60)
    lk[rng.random(5000) < 0.05] = np.nan
    rk[rng.random(4000) < 0.05] = np.nan
    _check([lk], [rk], how, keep)
    # all-unique right keys: the left columns are used as they are (left index None) in a left join
    _check([rng.integers(0, 1000, 3000)], [rng.permutation(1500)], how, keep)
    # integer sentinel keys are invalid too
    lk = rng.integers(0, 30, 2000).astype(np.int32)
    rk = rng.integers(0, 30, 1000).astype(np.int32)
    lk[::17] = np.iinfo(np.int32).min
    rk[::13] = np.iinfo(np.int32).min
    _check([lk], [rk], how, keep)


@pytest.mark.parametrize("how", ["left", "right", "inner", "outer"])
@pytest.mark.parametrize("keep", [(None, None), (None, "first"), (None, "last"), ("last", None)])
def test_merge_indices_multikey_and_strings(how, keep):
    if how == "outer" and keep[0] is not None:
        return
    rng = np.random.default_rng(abs(hash((how, keep, 2))) % 2**32)
    n1, n2 = 6000, 4500
    names = np.array([b"alpha", b"be", b"gamma_long", b"", b"zz"], dtype="S10")
    la, ra = rng.integers(0, 40, n1), rng.integers(10, 60, n2)
    ls, rs = names[rng.integers(0, 5, n1)], names[rng.integers(0, 5, n2)]
    _check([ls], [rs], how, keep)
    _check([la, ls], [ra, rs], how, keep)
    # a NaN in ANY column makes the tuple null for the join; the outer join's right-only rows with `keep` then follow
    # GROUP BY semantics (null only when ALL columns are null) -- rt_merge.py:1316-1334, 1460-1492
    lf, rf = np.round(rng.random(n1) * 8), np.round(rng.random(n2) * 8)
    lg, rg = np.round(rng.random(n1) * 3), np.round(rng.random(n2) * 3)
    for arr in (lf, rf, lg, rg):
        arr[rng.random(len(arr)) < 0.1] = np.nan
    _check([lf, lg], [rf, rg], how, keep)
    _check([la, lf], [ra, rf], how, keep)


def test_merge_indices_heavy_keys_and_device_mode():
    """One key with very many partners (the warp-cooperative copy of long runs), and CUDA tensors in -> CUDA tensors out."""
    import torch
    from riptable_b200 import merge

    rng = np.random.default_rng(3)
    lk = rng.integers(0, 4, 300)
    rk = np.r_[np.zeros(60_000, np.int64), rng.integers(1, 6, 5000)]
    rng.shuffle(rk)
    for how in ("left", "inner", "right", "outer"):
        _check([lk], [rk], how, (None, None))
    li, ri, k = merge.merge_indices([torch.from_numpy(lk).cuda()], [torch.from_numpy(rk).cuda()], "inner")
    assert li.is_cuda and ri.is_cuda
    el, er, _ = mo.merge_indices([lk], [rk], "inner")
    np.testing.assert_array_equal(li.cpu().numpy(), el)
    np.testing.assert_array_equal(ri.cpu().numpy(), er)
    # the indices really join: keys agree on every output row
    assert (lk[el] == rk[er]).all()
    with pytest.raises(TypeError):
        merge.merge_indices([lk], [rk.astype(np.float64)], "left")
    with pytest.raises(ValueError):
        merge.merge_indices([lk], [rk], "cross")


def test_merge_result_too_large_is_refused():
    """70 000 x 70 000 rows of one key = 4.9e9 result rows: past 2^32, so a 32-bit total would wrap to a small number."""
    from riptable_b200 import merge

    left = np.zeros(70_000, np.int32)
    right = np.zeros(70_000, np.int32)
    with pytest.raises(ValueError):
        merge.merge_indices([left], [right], how="inner")
