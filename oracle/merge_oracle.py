"""CPU restatement of riptable's join-index construction — TEST INFRASTRUCTURE, not product code.

Follows /root/reference/riptable/rt_merge.py step by step, in NumPy, on top of the oracle's grouping functions:
  merge_indices                 rt_merge.py:2080-2292  (keep handling, which side gets its NaN keys forced invalid)
  _get_or_create_keygroup       rt_merge.py:1934-2077  (valid-row masks; JOIN vs GROUP BY null semantics)
  _create_merge_fancy_indices   rt_merge.py:1285-1625  (left / right / inner / outer)
  _build_left_fancyindex        rt_merge.py:978-1190
  _build_right_fancyindex       rt_merge.py:734-884 with the numba loops :611-665 (_build_right_fancyindex_impl),
                                :668-731 (_fancy_index_fetch_keymap), :887-975 (_partial_repeat, _fused_arange_repeat)
The reference's merge code is executable Python / numba but cannot be imported here (riptable does not import on this
image's numpy), so it is restated; tests/test_oracle_golden.py pins this restatement to pandas.merge (row
order of left / right / inner joins) and to the reference's docstring examples.

Only non-Categorical key columns are handled (Categorical keys are container logic, out of scope).  Index dtypes are
int64 here; riptable narrows them by length (int_dtype_from_len), which the parity tests do not compare.
"""
import numpy as np

from . import riptide_oracle as orc

INV = np.iinfo(np.int64).min  # riptable's integer invalid for an int64 index


class KeyGroup:
    """The slice of riptable's Grouping the join uses (rt_grouping.py): base-1 ikey (0 = invalid row), first / last row
    per key, the packed layout, and the unique key rows."""

    def __init__(self, keycols, valid_mask):
        self.keycols = [np.asarray(c) for c in keycols]
        n = len(self.keycols[0])
        if n:
            self.ikey, self.ifirstkey, self.unique_count = orc.MultiKeyGroupBy32(self.keycols, 0, valid_mask, 2)
        else:
            self.ikey, self.ifirstkey, self.unique_count = np.zeros(0, np.int32), np.zeros(0, np.int32), 0
        self.ikey = self.ikey.astype(np.int64)
        u = self.unique_count
        self.ncountgroup, self.igroup, self.ifirstgroup = (x.astype(np.int64) for x in orc.BinCount(self.ikey, u + 1, pack=True))
        self.ifirstkey = self.ifirstkey.astype(np.int64)
        last = np.full(u + 1, -1, dtype=np.int64)
        last[self.ikey] = np.arange(n)  # later rows overwrite earlier ones
        self.ilastkey = last[1:]
        self.gbkeys = [c[self.ifirstkey] for c in self.keycols]  # _gbkeys_extract, rt_merge.py:1269-1282


def _valid_mask(col):
    """_create_column_valid_mask (rt_merge.py:1848-1872): strings have no invalid; isnotnan otherwise (NaN for floats,
    the dtype's sentinel for integers)."""
    col = np.asarray(col)
    if col.dtype.kind in "US":
        return None
    return ~orc._is_invalid(col)


def get_or_create_keygroup(keycols, force_invalids, create_groupby_grouping):
    """rt_merge.py:2000-2077.  -> (join KeyGroup, GROUP BY KeyGroup or None)."""
    join_mask, gb_mask, gb_all_valid = None, None, False
    if force_invalids:
        for col in keycols:
            v = _valid_mask(col)
            if v is not None:
                join_mask = v.copy() if join_mask is None else (join_mask & v)
            if create_groupby_grouping and len(keycols) > 1:
                if v is None:
                    gb_mask, gb_all_valid = None, True  # rt_merge.py:2056-2057: from here on the mask stays "all valid"...
                elif gb_mask is None:
                    gb_mask = v.copy()                  # ...unless a later column re-creates it (rt_merge.py:2058-2061)
                    gb_all_valid = False
                else:
                    gb_mask |= v
    join = KeyGroup(keycols, join_mask)
    groupby = KeyGroup(keycols, gb_mask) if (create_groupby_grouping and len(keycols) > 1) else None
    return join, groupby


def _keep_ifirstlast(g, keep):
    return None if keep is None else (g.ifirstkey if keep == "first" else g.ilastkey)


def _keep_ifirstlast_with_null(g, keep):
    """rt_merge.py:502-558: as above, with a leading entry for the invalid group when the data has invalid rows."""
    if keep is None:
        return None
    base = g.ifirstkey if keep == "first" else g.ilastkey
    if len(g.ncountgroup) > 0 and g.ncountgroup[0] > 0:
        extra = g.igroup[0] if keep == "first" else g.igroup[g.ncountgroup[0] - 1]
        return np.r_[extra, base]
    return base


def _ismember_keys(a_cols, b_cols):
    if len(a_cols) == 1:
        found, idx = orc.IsMember32(a_cols[0], b_cols[0])
    else:
        found, idx = orc.MultiKeyIsMember32((a_cols,), (b_cols,))
    return found, idx.astype(np.int64)


def _key_maps(left, right):
    """rt_merge.py:1342-1368: right key -> left key (0-based, ismember of the unique keys) and its inverse, 0-based right
    key id per left key with INV where the key does not occur in right."""
    if len(left.ikey) == 0 or len(right.ikey) == 0:
        exists = np.zeros(right.unique_count, dtype=bool)
        r2l = np.full(right.unique_count, INV, dtype=np.int64)
    else:
        exists, r2l = _ismember_keys(right.gbkeys, left.gbkeys)
    l2r = np.full(left.unique_count, INV, dtype=np.int64)
    l2r[r2l[exists]] = np.flatnonzero(exists)
    return exists, r2l, l2r


def _build_right_fancyindex(l2r, left, right, is_inner, keep):
    """rt_merge.py:734-884."""
    keep_left, keep_right = keep
    left_ikey = left.ikey if keep_left is None else left.ikey[_keep_ifirstlast_with_null(left, keep_left)]
    if len(left_ikey) == 0:
        return np.zeros(0, np.int64)
    if len(right.ikey) == 0:
        return np.zeros(0, np.int64) if is_inner else np.full(len(left_ikey), INV, dtype=np.int64)
    # _fancy_index_fetch_keymap (:668-731): 1-based right key id per left row, INV where none
    rows_keyid = np.full(len(left_ikey), INV, dtype=np.int64)
    ok = (left_ikey != 0) & (left_ikey <= len(l2r))
    cur = l2r[left_ikey[ok] - 1]
    rows_keyid[ok] = np.where(cur < 0, cur, cur + 1)
    if keep_right is not None:
        table = _keep_ifirstlast(right, keep_right)
        if is_inner:
            valid = rows_keyid != INV
            return table[rows_keyid[valid] - 1]
        out = np.full(len(rows_keyid), INV, dtype=np.int64)  # riptable's fancy index with an invalid yields the invalid
        valid = rows_keyid != INV
        out[valid] = table[rows_keyid[valid] - 1]
        return out
    counts = np.full(len(rows_keyid), 0 if is_inner else 1, dtype=np.int64)
    valid = rows_keyid != INV
    counts[valid] = right.ncountgroup[rows_keyid[valid]]
    cum = np.cumsum(counts)
    out = np.empty(int(cum[-1]) if len(cum) else 0, dtype=np.int64)
    # _build_right_fancyindex_impl (:611-665)
    group_count = len(right.ifirstgroup)
    for i in range(len(rows_keyid)):
        off = 0 if i == 0 else int(cum[i - 1])
        if off < len(out):
            keyid = rows_keyid[i]
            if keyid <= 0 or keyid >= group_count:
                if cum[i] > off:
                    out[off] = INV
            else:
                c = int(cum[i]) - off
                g0 = int(right.ifirstgroup[keyid])
                out[off:off + c] = right.igroup[g0:g0 + c]
    return out


def _build_left_fancyindex(r2l, exists, left, right, is_inner, keep):
    """rt_merge.py:978-1190.  None = the left columns are used as they are."""
    keep_left, keep_right = keep
    right_all_unique = right.unique_count == len(right.ikey)
    if keep_right is not None or right_all_unique:
        if is_inner:
            in_right = np.zeros(left.unique_count + 1, dtype=bool)
            in_right[1:][r2l[exists]] = True
            if keep_left is None:
                return np.flatnonzero(in_right[left.ikey])
            return _keep_ifirstlast(left, keep_left)[in_right[1:]]
        if keep_left is None:
            return None
        return _keep_ifirstlast_with_null(left, keep_left)
    mult = np.full(left.unique_count + 1, 0 if is_inner else 1, dtype=np.int64)
    mult[1:][r2l[exists]] = right.ncountgroup[1:][exists]
    left_ikey = left.ikey if keep_left is None else left.ikey[_keep_ifirstlast(left, keep_left)]
    repeats = mult[left_ikey]
    if keep_left is None:
        return np.repeat(np.arange(len(left_ikey), dtype=np.int64), repeats)  # _fused_arange_repeat (:933-975)
    return np.repeat(_keep_ifirstlast(left, keep_left), repeats)              # _partial_repeat (:887-930)


def _left_or_inner(left, right, is_inner, keep):
    exists, r2l, l2r = _key_maps(left, right)
    return (_build_left_fancyindex(r2l, exists, left, right, is_inner, keep),
            _build_right_fancyindex(l2r, left, right, is_inner, keep), None)


def _extract_groups(mask, igroup, ncountgroup, ifirstgroup):
    """Grouping.extract_groups (rt_grouping.py:3700-3790): the packed rows of the selected groups, group after group."""
    parts = [igroup[ifirstgroup[g]: ifirstgroup[g] + ncountgroup[g]] for g in np.flatnonzero(mask)]
    return np.concatenate(parts) if parts else np.zeros(0, np.int64)


def _outer(left, right, right_groupby, keep):
    """rt_merge.py:1404-1588."""
    exists, r2l, l2r = _key_maps(left, right)
    only = np.empty(right.unique_count + 1, dtype=bool)
    only[0] = right.ncountgroup[0] > 0
    only[1:] = ~exists
    if keep[1] is None:
        right_only_rows = _extract_groups(only, right.igroup, right.ncountgroup, right.ifirstgroup)
    elif right_groupby is not None and right.ncountgroup[0] > 0:
        inv_rows = right.igroup[: right.ncountgroup[0]]
        gb_groups = np.unique(right_groupby.ikey[inv_rows])
        table = _keep_ifirstlast_with_null(right_groupby, keep[1])
        # (the table has a leading invalid-group entry exactly when the GROUP BY grouping has invalid rows; group ids of
        # that grouping index it directly in that case, and shifted by one otherwise -- rt_merge.py:1478-1480 indexes it
        # directly, which is only right in the first case)
        has_null = len(right_groupby.ncountgroup) > 0 and right_groupby.ncountgroup[0] > 0
        inv_picked = table[gb_groups if has_null else gb_groups - 1]
        valid_picked = _keep_ifirstlast(right, keep[1])[only[1:]]
        right_only_rows = np.r_[inv_picked, valid_picked]
    else:
        mask = only if only[0] else only[1:]
        right_only_rows = _keep_ifirstlast_with_null(right, keep[1])[mask]
    right_only_rows = np.sort(right_only_rows)
    li = _build_left_fancyindex(r2l, exists, left, right, False, keep)
    ri = _build_right_fancyindex(l2r, left, right, False, keep)
    k = len(right_only_rows)
    if k:
        if li is None:
            li = np.arange(len(left.ikey), dtype=np.int64)
        li = np.r_[li, np.full(k, INV, dtype=np.int64)]
        if ri is None:
            ri = np.arange(len(right.ikey), dtype=np.int64)
        ri = np.r_[ri, right_only_rows]
    return li, ri, k


def merge_indices(left_keys, right_keys, how="left", keep=(None, None)):
    """rt.merge_indices on lists of key columns (rt_merge.py:2080-2292) -> (left_index or None, right_index or None,
    right_only_rowcount or None): fancy indices into the left / right rows, INV where a side has no row."""
    if isinstance(keep, str) or keep is None:
        keep = (keep, keep)
    if how == "outer" and keep[0] is not None:
        raise ValueError("Specifying a 'keep' value for the left Dataset (or both) with an outer merge is temporarily not "
                         "permitted due to a bug in the implementation.")  # rt_merge.py:2166-2170
    left, _ = get_or_create_keygroup(left_keys, how != "left", False)
    right, right_gb = get_or_create_keygroup(right_keys, how != "right", how == "outer")
    if how == "left":
        return _left_or_inner(left, right, False, keep)
    if how == "right":
        li, ri, _ = _left_or_inner(right, left, False, (keep[1], keep[0]))
        return ri, li, None
    if how == "inner":
        return _left_or_inner(left, right, True, keep)
    if how == "outer":
        if len(left.ikey) == 0 or len(right.ikey) == 0:
            sub = "left" if len(right.ikey) == 0 else "right"
            li, ri, _ = merge_indices(left_keys, right_keys, sub, keep)
            return li, ri, (0 if len(right.ikey) == 0 or ri is None else len(ri))
        return _outer(left, right, right_gb, keep)
    raise ValueError(f"Unsupported value ({how}) specified for the 'how' parameter.")
