"""CPU ORACLE (test infrastructure, NOT a product path).

A NumPy restatement of the riptide_cpp entry points on riptable's groupby / reduce / lexsort hot
path.  riptide_cpp itself (pinned >=1.19.0,<2 by /root/reference pyproject.toml:15) is a third-party
C++ package whose source is not vendored under /root/reference and which is not installable here,
so this file restates its *published behaviour* as riptable's own call sites, docstring golden
vectors and tests pin it.  Parity is PINNED against those vectors in tests/test_oracle_golden.py
(docstring examples rt_numpy.py:2013-2062, 1926-1933, 2126-2156, 1533-1540 and the KATs of
tests/test_groupby.py, test_pgroupby.py, test_lexsort.py, test_grouping.py).  Behaviour no reference
test fixes is marked UNPINNED here and in DESIGN.md.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  riptable_b200 never does.

Every function cites the reference file:line whose contract it follows.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------
# constants (riptable/rt_enum.py:486-532, 88-143)
GB_SUM, GB_MEAN, GB_MIN, GB_MAX, GB_VAR, GB_STD = 0, 1, 2, 3, 4, 5
GB_NANSUM, GB_NANMEAN, GB_NANMIN, GB_NANMAX, GB_NANVAR, GB_NANSTD = 50, 51, 52, 53, 54, 55
GB_FIRST, GB_NTH, GB_LAST = 100, 101, 102
GB_CUMSUM = 300


def invalid_for(dt):
    """riptable/rt_enum.py:116-143 (non-win32 table)."""
    dt = np.dtype(dt)
    if dt.kind == "f":
        return dt.type(np.nan)
    if dt.kind == "b":
        return np.bool_(False)
    if dt.kind == "i":
        return dt.type(np.iinfo(dt).min)
    if dt.kind == "u":
        return dt.type(np.iinfo(dt).max)
    if dt.kind == "S":
        return b""
    if dt.kind == "U":
        return ""
    raise TypeError(dt)


def _index_dtype(n):
    # LexSort32 vs LexSort64 split at 2e9 rows (rt_numpy.py:2199-2202, 2669-2671)
    return np.int64 if n > 2e9 else np.int32


def _is_valid_mask(v):
    """True where v is neither NaN (floats) nor the dtype sentinel (ints); bool has no invalid."""
    if v.dtype.kind == "f":
        return ~np.isnan(v)
    if v.dtype.kind in "iu":
        return v != invalid_for(v.dtype)
    return np.ones(len(v), dtype=bool)


# ----------------------------------------------------------------------------------------------
# a1  rc.MultiKeyGroupBy32  (rt_numpy.py:1979-2094, call :2073)
def _bytes_view(a):
    """Per-row byte matrix: key equality is BYTEWISE (tests/test_lexsort.py:225-256 NaNs form one
    group; tests/test_unique.py:14-48 denormal patterns are distinct).  -0.0 vs +0.0: UNPINNED,
    bytewise here so they are distinct keys."""
    a = np.ascontiguousarray(a)
    return a.view(np.uint8).reshape(len(a), a.dtype.itemsize)


def _row_codes(list_arrays):
    """1-D array whose == is bytewise tuple equality across all key columns."""
    if len(list_arrays) == 1:
        a = np.ascontiguousarray(list_arrays[0])
        if a.dtype.itemsize in (1, 2, 4, 8) and a.dtype.kind not in "SUV":
            return a.view(f"u{a.dtype.itemsize}")
    mats = [_bytes_view(a) for a in list_arrays]
    rec = np.ascontiguousarray(np.concatenate(mats, axis=1))
    w = rec.shape[1]
    if w <= 8:
        pad = np.zeros((rec.shape[0], 8), dtype=np.uint8)
        pad[:, :w] = rec
        return pad.view(np.uint64).ravel()
    return rec.view(np.dtype((np.void, w))).ravel()


def _first_appearance(codes):
    """codes -> (ikey 1-based int64, ifirst int64) with bins numbered by first appearance."""
    n = len(codes)
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    _, first_idx, inverse = np.unique(codes, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")  # sorted-unique id -> appearance order
    rank = np.empty(len(order), dtype=np.int64)
    rank[order] = np.arange(len(order), dtype=np.int64)
    return rank[inverse.ravel()] + 1, first_idx[order].astype(np.int64)


def MultiKeyGroupBy32(list_arrays, hint_size=0, filter=None, hash_mode=2, cutoffs=None):
    """-> (iKey int32[N] base-1, 0 = filtered; iFirstKey int32[U]; unique_count).
    With cutoffs: numbering restarts per partition, iFirstKey = (partition-relative first rows,
    cumulative unique counts)  (tests/test_pgroupby.py:7-86)."""
    if isinstance(list_arrays, np.ndarray):
        list_arrays = [list_arrays]
    n = len(list_arrays[0])
    for a in list_arrays:
        if len(a) != n:
            raise ValueError("groupbyhash all arrays must have same length")
    codes = _row_codes(list_arrays)
    if filter is not None:
        filter = np.asarray(filter, dtype=bool)
        if len(filter) != n:
            raise ValueError("filter length mismatch")

    def one(lo, hi):
        c = codes[lo:hi]
        if filter is None:
            ik, ifirst = _first_appearance(c)
            return ik, ifirst
        f = filter[lo:hi]
        idx = np.flatnonzero(f)
        ik_f, ifirst_f = _first_appearance(c[idx])
        ik = np.zeros(hi - lo, dtype=np.int64)
        ik[idx] = ik_f
        return ik, idx[ifirst_f] if len(ifirst_f) else ifirst_f

    if cutoffs is None:
        ik, ifirst = one(0, n)
        return ik.astype(np.int32), ifirst.astype(np.int32), int(len(ifirst))

    cutoffs = np.asarray(cutoffs, dtype=np.int64)
    ikey = np.zeros(n, dtype=np.int32)
    firsts, ucut, lo, total = [], [], 0, 0
    for hi in cutoffs:
        ik, ifirst = one(lo, int(hi))
        ikey[lo:hi] = ik
        firsts.append(ifirst.astype(np.int32))
        total += len(ifirst)
        ucut.append(total)
        lo = int(hi)
    vals = np.concatenate(firsts) if firsts else np.zeros(0, np.int32)
    return ikey, (vals, np.asarray(ucut, dtype=np.int64)), total


# ----------------------------------------------------------------------------------------------
# a4  rc.CombineFilter (rt_numpy.py:1489-1508): where(filter, key, 0), same dtype
def CombineFilter(key, filter):
    key = np.asarray(key)
    return np.where(np.asarray(filter, dtype=bool), key, key.dtype.type(0)).astype(key.dtype)


def CombineAccum1Filter(key1, unique_count1, filter=None):
    """rt_numpy.py:1512-1545: drop filtered rows and bin 0, renumber the surviving bins by first
    appearance; iKey keeps key1's dtype, iFirstKey is int32 (docstring :1533-1540)."""
    key1 = np.asarray(key1)
    keep = key1 > 0
    if filter is not None:
        keep &= np.asarray(filter, dtype=bool)
    idx = np.flatnonzero(keep)
    ik_f, ifirst_f = _first_appearance(key1[idx].astype(np.int64))
    out = np.zeros(len(key1), dtype=key1.dtype)
    out[idx] = ik_f
    return out, idx[ifirst_f].astype(np.int32), int(len(ifirst_f))


def CombineAccum2Filter(key1, key2, unique_count1, unique_count2, filter=None):
    """rt_numpy.py:1549-1597.  The cell grid is (unique_count1+1) x (unique_count2+1) including the
    invalid row/column; iKey = key2*(unique_count1+1) + key1 (UNPINNED layout beyond the size of
    nCountGroup and 'False -> 0'; follows cat2keys rt_numpy.py:1714-1734)."""
    key1 = np.asarray(key1).astype(np.int64)
    key2 = np.asarray(key2).astype(np.int64)
    ncell = (unique_count1 + 1) * (unique_count2 + 1)
    ik = key2 * (unique_count1 + 1) + key1
    if filter is not None:
        ik = np.where(np.asarray(filter, dtype=bool), ik, 0)
    dt = np.int8 if ncell < 100 else np.int16 if ncell < 30_000 else np.int32 if ncell < 2_000_000_000 else np.int64
    return ik.astype(dt), np.bincount(ik, minlength=ncell).astype(np.int32)


# ----------------------------------------------------------------------------------------------
# a6/a7  rc.BinCount / rc.GroupFromBinCount / rc.GroupByPack32 (rt_numpy.py:1901-1975)
def BinCount(ikey, n, pack=False):
    """histogram int32[n]; pack=True -> (nCountGroup, iGroup, iFirstGroup) (rt_numpy.py:1960)."""
    ikey = np.asarray(ikey)
    cnt = np.bincount(ikey.astype(np.int64), minlength=n)[:n].astype(np.int32)
    if not pack:
        return cnt
    igroup, ifirstgroup = GroupFromBinCount(ikey, cnt)
    return cnt, igroup, ifirstgroup


def GroupFromBinCount(ikey, ncountgroup):
    """iGroup = rows bucketed by bin (bin 0 first), ascending row order inside a bin (stable);
    iFirstGroup = exclusive cumsum of nCountGroup  (docstring rt_numpy.py:1926-1933)."""
    ikey = np.asarray(ikey)
    idt = _index_dtype(len(ikey))
    igroup = np.argsort(ikey, kind="stable").astype(idt)
    ifirst = np.zeros(len(ncountgroup), dtype=np.int64)
    ifirst[1:] = np.cumsum(ncountgroup.astype(np.int64))[:-1]
    return igroup, ifirst.astype(idt)


def _partitions(cutoffs, n):
    """Row ranges [lo, hi) of the `cutoffs` partitions (int64 cumulative row ends, rt_numpy.py:1999-2000)."""
    cut = np.asarray(cutoffs, dtype=np.int64)
    if cut.ndim != 1 or len(cut) == 0 or cut[-1] != n or (np.diff(cut) < 0).any() or cut[0] < 0:
        raise ValueError("cutoffs must be a non-decreasing 1-D int64 array whose last entry is the row count")
    return list(zip(np.r_[0, cut[:-1]].tolist(), cut.tolist()))


def GroupByPack32(ikey, _unused, n, cutoffs=None):
    """Old cutoffs-capable packer (rt_numpy.py:1972) -> (iGroup, iFirstGroup, nCountGroup).
    Without cutoffs it equals BinCount(pack=True).

    cutoffs: UNPINNED by the reference's tests, DERIVED from the one cutoffs case they do pin
    (tests/test_pgroupby.py:7-86: every partition behaves as an array of its own and the results are laid end to
    end).  Partition p holds rows [c[p-1], c[p]) whose iKey is numbered per partition (what groupbyhash(cutoffs=)
    produces); it is packed on its own with U_p + 1 bins, U_p = the largest iKey it holds:
    iGroup[c[p-1]:c[p]] = that partition's packed rows (partition-relative), and iFirstGroup / nCountGroup are the
    per-partition arrays (each with its own slot 0) laid end to end (length sum(U_p + 1))."""
    if cutoffs is None:
        cnt, igroup, ifirstgroup = BinCount(ikey, n, pack=True)
        return igroup, ifirstgroup, cnt
    ikey = np.asarray(ikey)
    ig, ifg, cnt = [], [], []
    for lo, hi in _partitions(cutoffs, len(ikey)):
        part = ikey[lo:hi]
        nb = int(part.max()) + 1 if hi > lo else 1
        c, g, f = BinCount(part, nb, pack=True)
        ig.append(g.astype(np.int32)); ifg.append(f.astype(np.int32)); cnt.append(c)
    return np.concatenate(ig), np.concatenate(ifg), np.concatenate(cnt)


# ----------------------------------------------------------------------------------------------
# a5  rc.GroupByAll32 (rt_numpy.py:1858-1871; dispatch rt_grouping.py:3395-3397, 3433-3436)
def _sum_out_dtype(dt):
    if dt.kind == "f":
        return dt
    if dt.kind == "u" and dt.itemsize == 8:
        return np.dtype(np.uint64)  # UNPINNED (rt_groupbynumba.py:642-648 intent)
    return np.dtype(np.int64)


def _reduce_one(v, ikey, nbins, func, lo, hi):
    dt = v.dtype
    if dt.kind not in "biuf":
        return None  # unsupported value dtype -> None slot (rt_grouping.py:2341-2345)
    ik = ikey.astype(np.int64)
    inrange = (ik >= lo) & (ik < hi)
    nan_op = func >= GB_NANSUM
    base = func - 50 if nan_op else func
    use = inrange & _is_valid_mask(v) if nan_op else inrange
    ikx, vx = ik[use], v[use]
    cnt = np.bincount(ikx, minlength=nbins).astype(np.int64)

    if base == GB_SUM:
        odt = _sum_out_dtype(dt)
        if dt.kind == "f":
            out = np.bincount(ikx, weights=vx.astype(np.float64), minlength=nbins)
            return out.astype(odt)
        out = np.zeros(nbins, dtype=odt)
        np.add.at(out, ikx, vx.astype(odt))  # wrap-around exact
        return out
    if base == GB_MEAN:
        # mean -> float64 for every input dtype (float32 output dtype UNPINNED)
        s = np.bincount(ikx, weights=vx.astype(np.float64), minlength=nbins)
        with np.errstate(invalid="ignore", divide="ignore"):
            out = s / cnt
        out[cnt == 0] = np.nan
        return out
    if base in (GB_MIN, GB_MAX):
        inv = invalid_for(dt)
        out = np.full(nbins, inv, dtype=dt)
        if len(vx) == 0:
            return out
        order = np.argsort(ikx, kind="stable")
        iks, vs = ikx[order], vx[order]
        starts = np.flatnonzero(np.r_[True, iks[1:] != iks[:-1]])
        ufn = np.minimum if base == GB_MIN else np.maximum  # NaN-propagating
        red = ufn.reduceat(vs, starts)
        if not nan_op and dt.kind in "iu":
            # integer sentinel propagates through min AND max (tests/test_groupbyops.py:389-456)
            has_inv = np.add.reduceat((vs == inv).astype(np.int64), starts) > 0
            red = np.where(has_inv, inv, red)
        out[iks[starts]] = red
        return out
    if base in (GB_VAR, GB_STD):
        x = vx.astype(np.float64)
        s = np.bincount(ikx, weights=x, minlength=nbins)
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = s / cnt
            d = x - mean[ikx]
            m2 = np.bincount(ikx, weights=d * d, minlength=nbins)
            out = m2 / (cnt - 1)  # ddof = 1 (tests/test_groupby.py:78-96)
        out[cnt < 2] = np.nan  # 1-element group -> NaN (tests/test_groupby.py:191-210)
        if base == GB_STD:
            out = np.sqrt(out)
        return out
    raise ValueError(f"GroupByAll32: unsupported function {func}")


def GroupByAll32(values, ikey, unique_rows, funcList, binLowList, binHighList, func_param=0):
    """-> tuple of arrays[unique_rows+1] (slot 0 = filtered bin), None for unsupported columns.
    Slots outside [binLow, binHigh) hold the op's empty value (sum 0, others invalid/NaN)."""
    ikey = np.asarray(ikey)
    out = []
    for v, f, lo, hi in zip(values, funcList, binLowList, binHighList):
        out.append(_reduce_one(np.asarray(v), ikey, unique_rows + 1, int(f), int(lo), int(hi)))
    return tuple(out)


# ----------------------------------------------------------------------------------------------
# a8  rc.GroupByAllPack32 (rt_numpy.py:1875-1891; rt_grouping.py:3442-3454)
# 8f.2  per-group quantiles: GB_MEDIAN / GB_QUANTILE_MULT (rt_enum.py:506-509; rt_groupbyops.py:1560-1745, 2498-2513)
GB_MEDIAN, GB_QUANTILE_MULT = 103, 106
QUANTILE_MULTIPLIER = 10**9


GB_MODE = 104


def _mode(v, ikey, nb, lo, hi):
    """GB_MODE "auto handles nan" (rt_enum.py:507): the most frequent valid value of each bin in the value's own dtype
    (tests/test_groupby.py:838-861: [1, 2, 2] -> 2).  UNPINNED: among equally frequent values the smallest wins (what
    pandas' mode()[0] gives); a bin without a valid value gets the invalid."""
    if v.dtype.kind not in "biuf":
        return None
    out = np.empty(nb, v.dtype)
    out[:] = invalid_for(v.dtype)
    ik = ikey.astype(np.int64)
    ik[(ik < 0) | (ik >= nb)] = 0
    bad = _is_invalid(v)
    order = np.argsort(ik, kind="stable")
    bounds = np.searchsorted(ik[order], np.arange(nb + 1))
    for b in range(max(lo, 0), min(hi, nb)):
        rows = order[bounds[b]: bounds[b + 1]]
        vals = v[rows][~bad[rows]]
        if len(vals) == 0:
            continue
        u, c = np.unique(vals, return_counts=True)
        out[b] = u[np.argmax(c)]  # first maximum = smallest value
    return out


def quantile_func_param(q, is_nan_function):
    """rt_groupbyops.py:1720-1731."""
    return int(q * 1e9) + int(bool(is_nan_function)) * (QUANTILE_MULTIPLIER + 1)


def _quantile(v, ikey, nb, f, func_param, lo, hi):
    """float64[nb]; numpy's "midpoint" rule written as (lower + higher) / 2 so that infinities behave
    (rt_numpy.py:4153-4245, tests/test_groupbyops.py:540-700).  NaN and integer sentinels are invalid:
    the nan version skips them, the plain one returns NaN for a group that holds one (the reference test
    compares integer columns through their float copy with NaN at the sentinels).  GB_MEDIAN
    "auto handles nan" (rt_enum.py:506).  Empty group, or bin outside [lo, hi) -> NaN."""
    if v.dtype.kind not in "biuf":
        return None
    if f == GB_MEDIAN:
        q, nan_aware = 0.5, True
    else:
        nan_aware = func_param >= QUANTILE_MULTIPLIER + 1
        q = (func_param - nan_aware * (QUANTILE_MULTIPLIER + 1)) / 1e9
    out = np.full(nb, np.nan)
    ik = ikey.astype(np.int64)
    ik[(ik < 0) | (ik >= nb)] = 0
    bad = _is_invalid(v)
    x = v.astype(np.float64)
    order = np.argsort(ik, kind="stable")
    bounds = np.searchsorted(ik[order], np.arange(nb + 1))
    for b in range(max(lo, 0), min(hi, nb)):
        rows = order[bounds[b]: bounds[b + 1]]
        if len(rows) == 0:
            continue
        ok = ~bad[rows]
        if not nan_aware and not ok.all():
            continue
        vals = np.sort(x[rows][ok])
        m = len(vals)
        if m == 0:
            continue
        idx = (m - 1) * q  # numpy's virtual index, in float64 and unsnapped as np.quantile leaves it
        with np.errstate(invalid="ignore"):
            out[b] = (vals[int(np.floor(idx))] + vals[int(np.ceil(idx))]) / 2
    return out


# 8f.2  rolling functions of rc.GroupByAllPack32 (rt_enum.py:513-519; rt_groupbyops.py:2941-3153, 3500-3670)
GB_ROLLING_SUM, GB_ROLLING_NANSUM, GB_ROLLING_DIFF, GB_ROLLING_SHIFT = 200, 201, 202, 203
GB_ROLLING_COUNT, GB_ROLLING_MEAN, GB_ROLLING_NANMEAN = 204, 205, 206


GB_ROLLING_QUANTILE = 207
ROLLING_QUANTILE_MULTIPLIER = 10**9


def rolling_quantile_func_param(q, window):
    """rt_utils.py:1172-1184."""
    return int(q * 1e9) + int(window) * (ROLLING_QUANTILE_MULTIPLIER + 1)


def _rolling_quantile(v, ikey, iGroup, iFirstGroup, nCountGroup, nb, func_param):
    """float64[nrows] in original row order: the nan-quantile (midpoint rule as (lower + higher) / 2,
    rt_numpy.py:4202-4254) of the last `window` rows of the row's group; NaN / integer sentinels are
    skipped, a window without a valid value gives NaN (tests/test_groupbyops.py:901-1110 compare the rows
    from the first full window on with np_rolling_nanquantile).  UNPINNED: rows before the first full
    window use the rows that are there, like the rolling sums."""
    import warnings
    if v.dtype.kind not in "biuf":
        return None
    window = func_param // (ROLLING_QUANTILE_MULTIPLIER + 1)
    q = (func_param % (ROLLING_QUANTILE_MULTIPLIER + 1)) / 1e9
    if window < 1:
        raise ValueError("rolling window must be >= 1")
    n = len(iGroup)
    rows = iGroup.astype(np.int64)
    x = v.astype(np.float64)
    x[_is_invalid(v)] = np.nan
    xs = x[rows]  # packed order
    b = ikey.astype(np.int64)[rows]
    b[(b < 0) | (b >= nb)] = 0
    out_packed = np.full(n, np.nan)
    bounds = np.flatnonzero(np.r_[True, b[1:] != b[:-1], True]) if n else np.zeros(1, np.int64)
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for s, e in zip(bounds[:-1], bounds[1:]):
            seg = xs[s:e]
            w = min(window, len(seg))
            padded = np.r_[np.full(w - 1, np.nan), seg]
            for c0 in range(0, len(seg), 4096):  # bounded memory: 4096 windows at a time
                win = np.lib.stride_tricks.sliding_window_view(padded[c0: c0 + 4096 + w - 1], w)
                lo = np.nanquantile(win, q, axis=-1, method="lower")
                hi = np.nanquantile(win, q, axis=-1, method="higher")
                out_packed[s + c0: s + c0 + len(lo)] = (lo + hi) / 2
    out = np.empty(n, np.float64)
    out[rows] = out_packed
    return out


def _rolling(v, ikey, iGroup, iFirstGroup, nCountGroup, nb, f, window):
    """Row-shaped rolling results in original row order.  Bin 0 is computed as a group of its own
    (rt_grouping.py:3324-3326: base_bin = 0) except for cumcount, whose filtered rows are invalid
    (tests/test_categorical_groupby.py:760-763).
    shift / diff: the value `window` places earlier in the group (later if negative); invalid where
    there is none (pandas groupby.shift / diff, tests/test_groupby.py:652-708).
    sum / nansum / mean / nanmean: over the last `window` rows of the group including this one;
    partial windows use what is there (tests/test_fastarray.py:1297-1305 pins this for the
    ungrouped rc.Rolling: nansum([1,2,3,nan,nan,nan], 3) == [1,3,6,5,3,0], also with int64
    sentinels); nan variants skip NaN / sentinels, nanmean of nothing is NaN.
    UNPINNED: output dtypes (sum as cumsum; mean float64; diff keeps the input dtype; cumcount int32)."""
    n = len(iGroup)
    rows = iGroup.astype(np.int64)
    b = ikey.astype(np.int64)[rows]
    b[(b < 0) | (b >= nb)] = 0
    start = iFirstGroup.astype(np.int64)[b]
    cnt = nCountGroup.astype(np.int64)[b]
    j = np.arange(n, dtype=np.int64)
    if f == GB_ROLLING_COUNT:
        p = j - start
        out = np.empty(n, np.int32)
        out[rows] = np.where(b == 0, np.iinfo(np.int32).min, p if window >= 0 else cnt - 1 - p)
        return out
    if f in (GB_ROLLING_SHIFT, GB_ROLLING_DIFF):
        if f == GB_ROLLING_DIFF and v.dtype.kind not in "iuf":
            return None
        q = j - window
        ok = (q >= start) & (q < start + cnt)
        out = np.empty(n, v.dtype)
        out[:] = invalid_for(v.dtype)
        src = v[rows[q[ok]]]
        if f == GB_ROLLING_SHIFT:
            out[rows[ok]] = src
        else:
            with np.errstate(over="ignore", invalid="ignore"):
                out[rows[ok]] = v[rows[ok]] - src
        return out
    if v.dtype.kind not in "biuf":
        return None
    if window < 1:
        raise ValueError("rolling window must be >= 1")
    skip = f in (GB_ROLLING_NANSUM, GB_ROLLING_NANMEAN)
    mean = f in (GB_ROLLING_MEAN, GB_ROLLING_NANMEAN)
    odt = np.dtype(np.float64) if mean else _sum_out_dtype(v.dtype)
    acc_dt = np.dtype(np.float64) if (mean or v.dtype.kind == "f") else np.dtype(odt)
    acc = np.zeros(n, acc_dt)
    used = np.zeros(n, np.int64)
    bad = _is_invalid(v)
    with np.errstate(over="ignore", invalid="ignore"):
        for back in range(min(window, int(cnt.max()) if n else 0) - 1, -1, -1):  # oldest -> newest
            q = j - back
            ok = q >= start
            src_rows = rows[np.where(ok, q, 0)]
            if skip:
                ok &= ~bad[src_rows]
            acc = np.where(ok, acc + v[src_rows].astype(acc_dt), acc)
            used += ok
        if mean:
            vals = np.where(used > 0, acc / np.maximum(used, 1), np.nan)
        else:
            vals = acc
    out = np.empty(n, odt)
    out[rows] = vals.astype(odt)
    return out


def GroupByAllPack32(
    values, ikey, iGroup, iFirstGroup, nCountGroup, unique_rows, funcList, binLowList, binHighList, func_param=0
):
    """first / nth / last through the packed layout; out of range or empty bin -> invalid
    (tests/test_groupby.py:137-159, 755-770)."""
    iGroup, iFirstGroup, nCountGroup = map(np.asarray, (iGroup, iFirstGroup, nCountGroup))
    nb = unique_rows + 1
    res = []
    for v, f, lo, hi in zip(values, funcList, binLowList, binHighList):
        v = np.asarray(v)
        f = int(f)
        if f == GB_ROLLING_QUANTILE:
            res.append(_rolling_quantile(v, np.asarray(ikey), iGroup, iFirstGroup, nCountGroup, nb, int(func_param)))
            continue
        if GB_ROLLING_SUM <= f <= GB_ROLLING_NANMEAN:
            res.append(_rolling(v, np.asarray(ikey), iGroup, iFirstGroup, nCountGroup, nb, f, int(func_param)))
            continue
        if f == GB_MODE:
            res.append(_mode(v, np.asarray(ikey), nb, int(lo), int(hi)))
            continue
        if f in (GB_MEDIAN, GB_QUANTILE_MULT):
            res.append(_quantile(v, np.asarray(ikey), nb, f, int(func_param), int(lo), int(hi)))
            continue
        if f not in (GB_FIRST, GB_NTH, GB_LAST):
            raise ValueError(f"GroupByAllPack32: unsupported function {f}")
        cnt = nCountGroup[:nb].astype(np.int64)
        first = iFirstGroup[:nb].astype(np.int64)
        if f == GB_FIRST:
            off = np.zeros(nb, np.int64)
        elif f == GB_LAST:
            off = cnt - 1
        else:
            n = int(func_param)
            off = np.full(nb, n, np.int64) if n >= 0 else cnt + n
        b = np.arange(nb)
        ok = (off >= 0) & (off < cnt) & (b >= lo) & (b < hi)
        out = np.zeros(nb, dtype=v.dtype)
        if v.dtype.kind not in "SUV":
            out[:] = invalid_for(v.dtype)
        out[ok] = v[iGroup[(first + off)[ok]]]
        res.append(out)
    return tuple(res)


# ----------------------------------------------------------------------------------------------
# a9 / 8f.2  rc.EmaAll32: GB_CUMSUM, GB_CUMPROD, GB_CUM[NAN]MIN / MAX
#   (rt_grouping.py:3427-3430; rt_groupbyops.py:3158-3258; function numbers rt_enum.py:523-532)
GB_CUMPROD, GB_CUMNANMAX, GB_CUMNANMIN, GB_CUMMAX, GB_CUMMIN = 302, 306, 307, 308, 309


def _is_invalid(v):
    """NaN for floats, the riptable sentinel for integers, nothing for bool (rt_enum.py INVALID_DICT)."""
    if v.dtype.kind == "f":
        return np.isnan(v)
    if v.dtype.kind in "iu":
        return v == invalid_for(v.dtype)
    return np.zeros(len(v), bool)


GB_EMADECAY, GB_EMANORMAL, GB_EMAWEIGHTED = 301, 304, 305


def _ema(values, ikey, funcNum, func_param):
    """rt_groupbyops.py:3306-3498, the docstring formulas walked row by row:
      decay:    Output[i] = Column[i] + LastEma[grp] * exp(-decay_rate * (Time[i] - LastTime[grp]))
      normal:   w = exp(-decay_rate * (Time[i] - LastTime[grp])); LastEma = Column[i] * (1 - w) + LastEma * w
      weighted: LastEma = Column[i] * (1 - decay_rate) + LastEma * decay_rate
    The docstring tables pin the start of a group: ema_normal / ema_weighted give the first value back
    (group2 == 1: 1.00, 2.50, 4.75 with decay 0.5), ema_decay starts from LastEma = 0.
    UNPINNED (this restatement's choice, documented in the C header): a row the op-filter excludes
    shows LastEma (0 before the first included row) and leaves LastTime alone; reset_filter restarts
    the group at that row; bin-0 rows get the invalid; output float32 for float32 input, else float64."""
    f = int(funcNum)
    rate, time, filt, reset = func_param[0], func_param[1], func_param[2], func_param[3]
    ik = np.asarray(ikey).astype(np.int64)
    n = len(ik)
    t = None if f == GB_EMAWEIGHTED else np.asarray(time).astype(np.float64)
    inc = ik > 0
    if filt is not None:
        inc &= np.asarray(filt, dtype=bool)
    rst = np.zeros(n, bool) if reset is None else np.asarray(reset, dtype=bool)
    res = []
    for v in values:
        v = np.asarray(v)
        if v.dtype.kind not in "biuf":
            res.append(None)
            continue
        x = v.astype(np.float64)
        odt = np.dtype(np.float32) if v.dtype == np.float32 else np.dtype(np.float64)
        out = np.full(n, np.nan)
        last, lastt, started = {}, {}, {}
        with np.errstate(over="ignore", invalid="ignore"):
            for i in range(n):
                g = ik[i]
                if g <= 0:
                    continue
                if not inc[i]:
                    out[i] = last.get(g, 0.0)
                    continue
                if rst[i] or not started.get(g, False):
                    y = x[i]
                elif f == GB_EMAWEIGHTED:
                    y = x[i] * (1.0 - rate) + last[g] * rate
                else:
                    w = np.exp(-rate * (t[i] - lastt[g]))
                    y = x[i] + last[g] * w if f == GB_EMADECAY else x[i] * (1.0 - w) + last[g] * w
                last[g], started[g] = y, True
                if t is not None:
                    lastt[g] = t[i]
                out[i] = y
        res.append(out.astype(odt))
    return res


def EmaAll32(values, ikey, unique_rows, funcNum, func_param):
    """Per-group running sum / product / min / max in original row order (tests/test_groupby.py:555-602,
    tests/test_groupbyops.py:458-538).
    func_param = (decay, time, filter, reset_filter).  Op-filter False -> the row gets the group's
    current running value; bin 0 row -> invalid of the OUTPUT dtype.
    cumsum / cumprod: output dtype ints -> int64 (uint64 stays), floats keep theirs (UNPINNED);
    integer sentinels are not skipped (tests/test_merge.py:6425-6426).
    cummin / cummax: output dtype == input dtype (test_groupbyops.py:524-525); GB_CUMNAN* skip NaN /
    sentinels like np.fmin/fmax.accumulate, GB_CUMMIN/MAX propagate them like np.minimum/maximum
    .accumulate (test_groupbyops.py:529-537); before a group's first valid value the result is invalid.
    reset_filter (UNPINNED): on a row that passes the filter and has reset set, restart the
    accumulator before applying the row (rt_fastarraynumba.py:506-524)."""
    f = int(funcNum)
    if f in (GB_EMADECAY, GB_EMANORMAL, GB_EMAWEIGHTED):
        return _ema(values, ikey, f, func_param)
    if f not in (GB_CUMSUM, GB_CUMPROD, GB_CUMNANMAX, GB_CUMNANMIN, GB_CUMMAX, GB_CUMMIN):
        raise ValueError(f"EmaAll32: unsupported function {funcNum}")
    filt = reset = None
    if isinstance(func_param, (tuple, list)) and len(func_param) >= 4:
        filt, reset = func_param[2], func_param[3]
    ik = np.asarray(ikey).astype(np.int64)
    n = len(ik)
    active = ik > 0
    add = active & np.asarray(filt, dtype=bool) if filt is not None else active
    order = np.argsort(ik, kind="stable")
    iks = ik[order]
    seg_start = np.r_[True, iks[1:] != iks[:-1]] if n else np.zeros(0, bool)
    if reset is not None:
        seg_start = seg_start | (np.asarray(reset, dtype=bool) & add)[order]
    start_idx = np.flatnonzero(seg_start)
    ends = np.r_[start_idx[1:], n] if n else np.zeros(0, np.int64)
    adds = add[order]
    res = []
    for v in values:
        v = np.asarray(v)
        if v.dtype.kind not in "biuf":
            res.append(None)
            continue
        if f in (GB_CUMSUM, GB_CUMPROD):
            odt = _sum_out_dtype(v.dtype)
            acc_dt = np.dtype(np.float64) if v.dtype.kind == "f" else np.dtype(odt)
            ident = acc_dt.type(0 if f == GB_CUMSUM else 1)
            xs = np.where(adds, v.astype(acc_dt)[order], ident)
            run = np.empty(n, acc_dt)
            acc = np.cumsum if f == GB_CUMSUM else np.cumprod
            with np.errstate(over="ignore", invalid="ignore"):
                # exact per-segment sequential accumulation (no cancellation against earlier segments)
                for s, e in zip(start_idx, ends):
                    run[s:e] = acc(xs[s:e])
                out = np.empty(n, dtype=odt)
                out[order] = run.astype(odt)
        else:
            odt = v.dtype
            ismin = f in (GB_CUMMIN, GB_CUMNANMIN)
            skipna = f in (GB_CUMNANMIN, GB_CUMNANMAX)
            vs = v[order]
            bad = _is_invalid(vs) & adds
            good = adds & ~bad
            if v.dtype.kind == "f":
                fill = v.dtype.type(np.inf if ismin else -np.inf)
            elif v.dtype.kind == "b":
                fill = np.bool_(ismin)
            else:
                fill = v.dtype.type(np.iinfo(v.dtype).max if ismin else np.iinfo(v.dtype).min)
            cand = np.where(good, vs, fill)
            run = np.empty(n, v.dtype)
            ok = np.empty(n, bool)
            acc = np.minimum.accumulate if ismin else np.maximum.accumulate
            for s, e in zip(start_idx, ends):
                run[s:e] = acc(cand[s:e])
                ok[s:e] = np.logical_or.accumulate(good[s:e])
                if not skipna:
                    ok[s:e] &= ~np.logical_or.accumulate(bad[s:e])
            run[~ok] = invalid_for(odt)
            out = np.empty(n, dtype=odt)
            out[order] = run
        out[~active] = invalid_for(odt)
        res.append(out)
    return res


# ----------------------------------------------------------------------------------------------
# a10  rc.MakeiFirst / rc.MakeiNext (rt_numpy.py:1767-1854; tests/test_grouping.py:210-296)
def MakeiFirst(key, n, filter=None, mode=0):
    """Array indexed by bin, slots 0..n, first (mode 0) / last (mode 1) row of each bin; invalid
    sentinel where a bin has no row; slot 0 is always invalid.  Row ids need an index dtype:
    int32 (tests assert ilast == 174762 for int8 keys, so it cannot be the key's dtype)."""
    key = np.asarray(key).astype(np.int64)
    idt = _index_dtype(len(key))
    out = np.full(n + 1, invalid_for(idt), dtype=idt)
    rows = np.arange(len(key), dtype=np.int64)
    ok = (key > 0) & (key <= n)
    if filter is not None:
        ok &= np.asarray(filter, dtype=bool)
    k, r = key[ok], rows[ok]
    if len(k):
        if mode == 0:
            out[k[::-1]] = r[::-1]
        else:
            out[k] = r
    return out


def MakeiNext(key, n, mode=0):
    """next (mode 0) / previous (mode 1) row with the same bin, invalid sentinel where none.
    Bin-0 rows get the sentinel (UNPINNED)."""
    key = np.asarray(key).astype(np.int64)
    N = len(key)
    idt = _index_dtype(N)
    out = np.full(N, invalid_for(idt), dtype=idt)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    same = ks[1:] == ks[:-1]
    if mode == 0:
        src, dst = order[:-1][same], order[1:][same]
    else:
        src, dst = order[1:][same], order[:-1][same]
    out[src] = dst
    out[key <= 0] = invalid_for(idt)
    return out


# a13  rc.ReverseShuffle (rt_grouping.py:1659; tests/test_lexsort.py:325-332)
def ReverseShuffle(perm):
    perm = np.asarray(perm)
    out = np.empty_like(perm)
    out[perm] = np.arange(len(perm), dtype=perm.dtype)
    return out


# ----------------------------------------------------------------------------------------------
# a11  rc.LexSort32/64 (rt_numpy.py:2658-2673; tests/test_lexsort.py:21-103)
def _descending_code(a):
    """Rank codes whose ascending order is a's descending order with NaN still last and ties kept
    in original row order (pandas sort_values(ascending=False) semantics, tests/test_dataset.py:1579-1595)."""
    _, inv = np.unique(a, return_inverse=True)  # NaNs sort last and collapse to one code
    inv = inv.ravel().astype(np.int64)
    top = inv.max() if len(inv) else 0
    code = top - inv
    if a.dtype.kind == "f":
        nan = np.isnan(a)
        if nan.any():
            code = np.where(nan, top + 1, top - 1 - inv)
    return code


def LexSort32(list_arrays, cutoffs=None, index=None, ascending=True):
    """Stable lexicographic argsort == np.lexsort: LAST array is the primary key, NaN last."""
    if isinstance(list_arrays, np.ndarray):
        list_arrays = [list_arrays]
    list_arrays = [np.asarray(a) for a in list_arrays]
    n = len(list_arrays[0])
    idt = _index_dtype(n)
    if isinstance(ascending, (bool, np.bool_)):
        asc = [bool(ascending)] * len(list_arrays)
    else:
        asc = [bool(x) for x in ascending]
        if len(asc) != len(list_arrays):
            raise ValueError("ascending must have one entry per key")
    if cutoffs is not None:
        # UNPINNED by the reference's tests, DERIVED from the pinned hash case (tests/test_pgroupby.py:7-86): every
        # partition is sorted as an array of its own; out[c[p-1]:c[p]] holds partition-relative row numbers.
        if index is not None:
            raise ValueError("lexsort: cutoffs and index (a filter) cannot be combined")
        parts = [LexSort32([k[lo:hi] for k in list_arrays], None, None, asc) for lo, hi in _partitions(cutoffs, n)]
        return np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, np.int32)
    keys = list_arrays
    if index is not None:
        index = np.asarray(index).astype(np.int64)
        keys = [k[index] for k in keys]
    keys = [k if a else _descending_code(k) for k, a in zip(keys, asc)]
    perm = np.lexsort(keys)
    if index is not None:
        perm = index[perm]
    return perm.astype(idt)


LexSort64 = LexSort32


# a12  rc.GroupFromLexSort (rt_numpy.py:2098-2251, call :2207)
def GroupFromLexSort(iGroup, value_array, cutoffs=None, base_index=1):
    """Walk the sorted order; a new bin starts where the key row differs (bytewise) from the
    previous one.  -> (iKey int32[N] in base `base_index`, iFirstKey[U] = first (smallest) row of
    each bin, nCountGroup int32[U+1] with slot 0 = 0).  Rows absent from iGroup keep iKey 0."""
    iGroup = np.asarray(iGroup)
    cols = value_array if isinstance(value_array, (list, tuple)) else [value_array]
    n = len(cols[0])
    if cutoffs is not None:
        # UNPINNED, DERIVED as for LexSort32: per-partition calls laid end to end.  iKey numbering restarts in every
        # partition, iFirstKey holds partition-relative rows, nCountGroup = [0] + every partition's counts, and the
        # fourth result is the cumulative unique count per partition (what groupbylex stores as nUniqueCutoffs,
        # rt_numpy.py:2209-2212; the shape test_pgroupby.py pins for iFirstKey[1]).
        if len(iGroup) != n:
            raise ValueError("GroupFromLexSort: cutoffs need an iGroup that covers every row (no filter)")
        ik, ifk, cnt, ucut = [], [], [np.zeros(1, np.int32)], []
        total = 0
        for lo, hi in _partitions(cutoffs, n):
            a, b, c = GroupFromLexSort(iGroup[lo:hi], [np.asarray(x)[lo:hi] for x in cols], None, base_index)
            ik.append(a.astype(np.int32)); ifk.append(b.astype(np.int32)); cnt.append(c[1:])
            total += len(b)
            ucut.append(total)
        return (np.concatenate(ik), np.concatenate(ifk), np.concatenate(cnt), np.asarray(ucut, dtype=np.int64))
    codes = _row_codes([np.asarray(c) for c in cols])
    idt = _index_dtype(n)
    ikey = np.zeros(n, dtype=idt)
    m = len(iGroup)
    if m == 0:
        return ikey, np.zeros(0, idt), np.zeros(1, np.int32)
    sc = codes[iGroup]
    newbin = np.r_[True, sc[1:] != sc[:-1]]
    binid = np.cumsum(newbin)  # 1-based
    ikey[iGroup] = binid - (1 - base_index)
    starts = np.flatnonzero(newbin)
    ifirstkey = iGroup[starts].astype(idt)
    ncount = np.zeros(len(starts) + 1, dtype=np.int32)
    ncount[1:] = np.diff(np.r_[starts, m])
    return ikey, ifirstkey, ncount


# ----------------------------------------------------------------------------------------------
# Python-level wrappers restated from rt_numpy.py so golden docstring vectors can be replayed
def groupbyhash(list_arrays, hint_size=0, filter=None, hash_mode=2, cutoffs=None, pack=False):
    """rt_numpy.py:1979-2094."""
    if isinstance(list_arrays, np.ndarray):
        list_arrays = [list_arrays]
    if len(list_arrays[0]) == 0:
        e = np.zeros(0, np.int32)
        d = {"iKey": e, "iFirstKey": e, "unique_count": 0}
    else:
        ik, ifk, u = MultiKeyGroupBy32(list_arrays, hint_size, filter, hash_mode, cutoffs=cutoffs)
        d = {"iKey": ik, "iFirstKey": ifk, "unique_count": u}
    if pack:
        d.update(groupbypack(d["iKey"], None, d["unique_count"] + 1))
    else:
        d.update({"iGroup": None, "iFirstGroup": None, "nCountGroup": None})
    return d


def groupbypack(ikey, ncountgroup, unique_count=None, cutoffs=None):
    """rt_numpy.py:1901-1975."""
    if cutoffs is not None:  # rt_numpy.py:1972
        igroup, ifirstgroup, cnt = GroupByPack32(ikey, None, unique_count, cutoffs=cutoffs)
        return {"iGroup": igroup, "iFirstGroup": ifirstgroup, "nCountGroup": cnt}
    if ncountgroup is None:
        cnt, igroup, ifirstgroup = BinCount(ikey, unique_count, pack=True)
    else:
        igroup, ifirstgroup = GroupFromBinCount(ikey, ncountgroup)
        cnt = ncountgroup
    return {"iGroup": igroup, "iFirstGroup": ifirstgroup, "nCountGroup": cnt}


def groupbylex(list_arrays, filter=None, cutoffs=None, base_index=1, rec=False):
    """rt_numpy.py:2098-2251 (Python glue restated around LexSort32 + GroupFromLexSort)."""
    if isinstance(list_arrays, np.ndarray):
        list_arrays = [list_arrays]
    value_cols = list(list_arrays)
    sort_keys = list_arrays[::-1] if len(list_arrays) > 1 else list_arrays
    index = None
    filterfalse = None
    if filter is not None:
        filter = np.asarray(filter, dtype=bool)
        index = np.flatnonzero(filter).astype(np.int32)
        filterfalse = np.flatnonzero(~filter)
    iGroup = LexSort32(sort_keys, cutoffs=cutoffs, index=index)
    retval = GroupFromLexSort(iGroup, value_cols, cutoffs=cutoffs, base_index=base_index)
    iKey, iFirstKey, nCountGroup = retval[:3]
    if base_index == 0:
        iKey = iKey + 1
    else:
        nCountGroup[0] = 0
    iFirstGroup = nCountGroup.copy()
    iFirstGroup[1:] = nCountGroup.cumsum()[:-1]
    if filter is not None:
        nCountGroup[0] = len(filterfalse)
        iKey[filterfalse] = 0
        iFirstGroup[0] = len(index)
    d = {
        "iKey": iKey,
        "iFirstKey": iFirstKey,
        "unique_count": len(iFirstKey),
        "iGroup": iGroup,
        "iFirstGroup": iFirstGroup,
        "nCountGroup": nCountGroup,
    }
    if len(retval) == 4:  # rt_numpy.py:2209-2212, 2249-2250
        d["nUniqueCutoffs"] = retval[3]
    return d


# ----------------------------------------------------------------------------------------------
# 8f.1  rc.IsMember32 / rc.MultiKeyIsMember32 (rt_numpy.py:1186-1393; tests/test_ismember.py)
def _index_dtype_from_len(n):
    """rt_enum.py:666-679 (tests/test_ismember.py:288-300: the index array is downcast by len(b))."""
    return np.int8 if n < 100 else np.int16 if n < 30_000 else np.int32 if n < 2_000_000_000 else np.int64


def _ismember_cols(a_cols, b_cols):
    a_cols = [np.asarray(c) for c in a_cols]
    b_cols = [np.asarray(c) for c in b_cols]
    na, nb = len(a_cols[0]), len(b_cols[0])
    idt = np.dtype(_index_dtype_from_len(nb))
    inv = invalid_for(idt)
    out = np.full(na, inv, dtype=idt)
    found = np.zeros(na, dtype=bool)
    if na == 0 or nb == 0:
        return found, out
    cat, valid = [], np.ones(na + nb, dtype=bool)
    for ac, bc in zip(a_cols, b_cols):
        if ac.dtype.kind in "SU" or bc.dtype.kind in "SU":
            w = max(ac.dtype.itemsize, bc.dtype.itemsize)
            kind = ac.dtype.kind
            dt = np.dtype(f"{kind}{w // (4 if kind == 'U' else 1)}")
            ac, bc = ac.astype(dt), bc.astype(dt)
        elif ac.dtype != bc.dtype:
            raise TypeError("ismember: numeric keys must share a dtype (riptable casts before calling rc)")
        c = np.concatenate([bc, ac])
        if c.dtype.kind == "f":
            valid &= ~np.isnan(c)  # NaN never matches (tests/test_ismember.py)
        cat.append(c)
    ik, ifk, _ = MultiKeyGroupBy32(cat, 0, valid, 2)
    ika = ik[nb:].astype(np.int64)
    first = ifk[np.maximum(ika - 1, 0)] if len(ifk) else np.zeros(na, np.int64)
    found = (ika > 0) & (first < nb)
    out[found] = first[found].astype(idt)
    return found, out


def IsMember32(a, b, h=2, hint_size=0):
    """-> (bool[len(a)], index of the first occurrence of a[i] in b, the index dtype's invalid where absent)."""
    return _ismember_cols([a], [b])


def IsMemberCategorical(a, b, h=2, hint_size=0):
    """rt_numpy.py:1388-1389 (`ismember(a, b, base_index=1)`): 1-based position of the first occurrence in b, 0 where
    absent -- the shape rt_grouping.py:959-963 feeds straight back in as Categorical codes.  UNPINNED: the index dtype
    (here int_dtype_from_len(len(b) + 1))."""
    found, idx = _ismember_cols([a], [b])
    odt = np.dtype(_index_dtype_from_len(len(np.asarray(b)) + 1))
    out = np.zeros(len(idx), odt)
    out[found] = idx[found].astype(np.int64) + 1
    return found, out


def IsMemberCategoricalFixup(a_codes, b_codes, on_unique, num_unique_b, a_base_index=1, b_base_index=1):
    """rc.IsMemberCategoricalFixup (rt_numpy.py:1288-1305): ismember of two Categoricals computed from the match of
    their uniques: code of a -> unique of a -> unique of b (`on_unique`, the index result of ismember(acats, bcats)) ->
    code of b -> FIRST row of b holding that code.  Equals ismember(a.as_string_array, b.as_string_array)
    (tests/test_ismember.py:162-195); filtered rows (code 0 in base 1) and sentinel codes match nothing
    (tests/test_ismember.py:491-505).  The index dtype is int32 (the tests cast it before comparing)."""
    a = np.asarray(a_codes).astype(np.int64)
    b = np.asarray(b_codes).astype(np.int64)
    on_unique = np.asarray(on_unique).astype(np.int64)
    nslots = int(num_unique_b) + int(b_base_index)
    first = np.full(nslots, -1, dtype=np.int64)
    okb = (b >= 0) & (b < nslots)
    rows = np.flatnonzero(okb)
    first[b[rows][::-1]] = rows[::-1]
    ua = a - int(a_base_index)
    inv = invalid_for(np.int32)
    found = np.zeros(len(a), dtype=bool)
    idx = np.full(len(a), inv, dtype=np.int32)
    ok = (ua >= 0) & (ua < len(on_unique))
    ub = np.where(ok, on_unique[np.clip(ua, 0, max(len(on_unique) - 1, 0))] if len(on_unique) else -1, -1)
    ok &= (ub >= 0) & (ub < int(num_unique_b))
    pos = np.where(ok, first[np.clip(ub + int(b_base_index), 0, max(nslots - 1, 0))] if nslots else -1, -1)
    found = pos >= 0
    idx[found] = pos[found]
    return found, idx


def MultiKeyIsMember32(a_tuple, b_tuple, hint_size=0):
    """rt_numpy.py:1325: rc.MultiKeyIsMember32((a_cols,), (b_cols,), hint_size) -- tuple equality across the columns."""
    return _ismember_cols(list(a_tuple[0]), list(b_tuple[0]))


# ----------------------------------------------------------------------------------------------
# 8f.3  rc.ReIndexGroups (rt_grouping.py:378; restates the Python fallback at rt_grouping.py:380-407, base_index 1)
def ReIndexGroups(ikey, uikey, u_cutoffs, i_cutoffs):
    """Per partition p: newikey[rows of p] = uikey[entries of p][old - 1]; a 0 stays 0."""
    ikey = np.asarray(ikey)
    uikey = np.asarray(uikey)
    new = np.zeros(len(ikey), dtype=ikey.dtype)
    start = starti = 0
    for stop, stopi in zip(np.asarray(u_cutoffs, np.int64), np.asarray(i_cutoffs, np.int64)):
        us = uikey[start:stop]
        old = ikey[starti:stopi].astype(np.int64)
        ok = (old > 0) & (old <= len(us))
        seg = np.zeros(len(old), dtype=ikey.dtype)
        seg[ok] = us[old[ok] - 1]
        new[starti:stopi] = seg
        start, starti = stop, stopi
    return new
