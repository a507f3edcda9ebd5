"""Join-index construction on the GPU: the device mirror of `rt.merge_indices` (riptable/rt_merge.py:2080-2292).

riptable turns a join into two fancy indices (one per Dataset) built from the two sides' groupings; `merge2`, `merge`
and `merge_lookup` then index every column with them (riptable/rt_merge.py:2293-2870).  The groupings come from the hot
path this package already provides (`rc.MultiKeyGroupBy32`, the packed layout, `ismember` of the unique keys); what
riptable adds on top in Python / numba — key-id mapping, per-row match counts, the cumulative sum and the two row
expansions — is done here with three CUDA passes (csrc/merge.cu).  Everything row-sized stays on the device; the
per-key tables (U-sized) are glued with torch ops.

    left_index, right_index, right_only_rowcount = merge_indices(left_keys, right_keys, how="left", keep=None)

`left_keys` / `right_keys`: lists of key columns (NumPy arrays in -> NumPy indices out; CUDA tensors / rc.DevArray in ->
CUDA tensors out).  Indices are int32 row numbers, the int32 invalid (INT32_MIN) where that side has no row; `None` means
"use that side's columns as they are" exactly where riptable returns None (rt_merge.py:1083-1090).  Semantics follow the
reference line by line (SQL JOIN null rules, `keep`, row order); oracle/merge_oracle.py is the CPU restatement the tests
compare against.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import rc
from ._lib import check, lib
from .enums import dtype_code

INV = np.iinfo(np.int32).min


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _i32(n, dev):
    return torch.empty(max(int(n), 4), dtype=torch.int32, device=dev)[: int(n)]


class _KeyGroup:
    """What the join needs of riptable's Grouping (rt_grouping.py): base-1 iKey with 0 = invalid row, first / last row
    per key, the packed layout (lazily), the unique key rows."""

    def __init__(self, cols, valid):
        self.cols = cols
        self.n = cols[0].n
        self.dev = rc._device()
        if self.n:
            ik, ifk, u = rc.MultiKeyGroupBy32(cols, 0, valid, 2)
            self.ikey, self.ifirstkey, self.u = ik, ifk.contiguous(), int(u)
        else:
            self.ikey, self.ifirstkey, self.u = _i32(0, self.dev), _i32(0, self.dev), 0
        self._packed = None
        self._ilast = None

    def packed(self):
        """(nCountGroup, iGroup, iFirstGroup), slot 0 = the invalid rows (rc.BinCount(pack=True), rt_numpy.py:1960)."""
        if self._packed is None:
            if self.n:
                self._packed = tuple(t.contiguous() for t in rc.BinCount(self.ikey, self.u + 1, pack=True))
            else:
                z = torch.zeros(1, dtype=torch.int32, device=self.dev)
                self._packed = (z, _i32(0, self.dev), z.clone())
        return self._packed

    def ilastkey(self):
        if self._ilast is None:
            self._ilast = rc.MakeiFirst(self.ikey, self.u, None, 1)[1:].contiguous() if self.n else _i32(0, self.dev)
        return self._ilast

    def keep_rows(self, keep, with_null):
        """ifirstkey / ilastkey, optionally with a leading entry for the invalid group (rt_merge.py:488-558)."""
        base = self.ifirstkey if keep == "first" else self.ilastkey()
        if with_null:
            cnt, ig, _ = self.packed()
            n0 = int(cnt[0].item()) if cnt.numel() else 0
            if n0 > 0:
                extra = ig[0:1] if keep == "first" else ig[n0 - 1: n0]
                return torch.cat([extra, base]).contiguous()
        return base

    def gbkeys(self):
        """The unique key rows, one DevArray per key column (`_gbkeys_extract`, rt_merge.py:1269-1282)."""
        out = []
        idx = self.ifirstkey.to(torch.int64)
        for c in self.cols:
            isz = c.dtype.itemsize
            rows = c.buf[: c.n * isz].view(c.n, isz).index_select(0, idx).reshape(-1)
            if rows.numel() < 16:
                pad = torch.zeros(16, dtype=torch.uint8, device=rows.device)
                pad[: rows.numel()] = rows
                rows = pad
            out.append(rc.DevArray(rows.contiguous(), c.dtype, self.u))
        return out


def _valid_mask(cols, how_or):
    """AND (JOIN semantics) or OR (GROUP BY semantics) of the columns' validity (rt_merge.py:2038-2063); None = all valid."""
    n = cols[0].n
    mask, all_valid = None, False
    for c in cols:
        if c.dtype.kind in "SUV":
            if how_or:
                mask, all_valid = None, True
            continue
        if c.dtype.kind not in "biuf" or c.dtype == np.float16:
            continue
        if how_or and all_valid and mask is None:
            all_valid = False  # rt_merge.py:2058-2061: a later column re-creates the mask
        mode = 0 if mask is None else (2 if how_or else 1)
        if mask is None:
            mask = torch.empty(max(n, 4), dtype=torch.uint8, device=rc._device())[:n]
        check(lib.rtb_valid_mask(C.c_void_p(c.ptr), dtype_code(c.dtype), n, mode, _vp(mask), rc._stream()))
    return None if mask is None else mask.view(torch.bool)


def _keygroups(cols, force_invalids, want_groupby):
    join = _KeyGroup(cols, _valid_mask(cols, False) if force_invalids else None)
    gb = None
    if want_groupby and len(cols) > 1:
        gb = _KeyGroup(cols, _valid_mask(cols, True) if force_invalids else None)
    return join, gb


def _key_maps(left: _KeyGroup, right: _KeyGroup):
    """rt_merge.py:1342-1368 -> (exists bool[U_r], r2l int64[U_r] 0-based left key or < 0, l2r1 int32[U_l] 1-based right
    key id or INV)."""
    dev = left.dev
    if left.n == 0 or right.n == 0 or left.u == 0 or right.u == 0:
        exists = torch.zeros(right.u, dtype=torch.bool, device=dev)
        r2l = torch.full((right.u,), -1, dtype=torch.int64, device=dev)
    else:
        a, b = right.gbkeys(), left.gbkeys()
        exists, idx = rc.IsMember32(a[0], b[0]) if len(a) == 1 else rc.MultiKeyIsMember32((a,), (b,))
        r2l = torch.where(exists, idx.to(torch.int64), torch.full_like(idx, -1, dtype=torch.int64))
    l2r1 = torch.full((max(left.u, 1),), INV, dtype=torch.int32, device=dev)[: left.u]
    sel = torch.nonzero(exists).reshape(-1)
    if sel.numel():
        l2r1[r2l[sel]] = (sel + 1).to(torch.int32)
    return exists, r2l, l2r1.contiguous()


def _expand(left_ikey, left_rows, keymap1, r_ncount, r_ifirstgroup, r_igroup, nonmatched, want_left, want_right, invalid_rows=None):
    """The three row-sized passes: match counts per driving row, their exclusive scan, the expansion.
    nonmatched: rows a driving row without a partner contributes (0 inner, 1 left join); invalid_rows: the same for a
    driving row whose own key is invalid (defaults to nonmatched)."""
    dev = left_ikey.device
    n = int(left_ikey.numel())
    if n == 0:
        return (_i32(0, dev) if want_left else None), (_i32(0, dev) if want_right else None)
    left_ikey = left_ikey.contiguous()
    keyid = _i32(n, dev)
    counts = _i32(n, dev)
    check(lib.rtb_merge_row_counts(_vp(left_ikey), n, _vp(keymap1), int(keymap1.numel()), _vp(r_ncount), int(r_ncount.numel()),
                                   int(nonmatched), int(nonmatched if invalid_rows is None else invalid_rows), _vp(keyid),
                                   _vp(counts), rc._stream()))
    offsets = _i32(n, dev)
    ws = rc._ws(lib.rtb_merge_scan_workspace_bytes(n))
    total = C.c_int64(0)
    check(lib.rtb_merge_scan_counts(_vp(counts), n, _vp(offsets), C.byref(total), C.c_void_p(ws.data_ptr()), ws.numel(), rc._stream()))
    m = int(total.value)
    if m > 2_147_483_646:
        raise ValueError("merge: the join has more than 2^31 - 2 rows")
    lo = _i32(m, dev) if want_left else None
    ro = _i32(m, dev) if want_right else None
    if m:
        check(lib.rtb_merge_expand(_vp(keyid), _vp(offsets), _vp(counts), n, _vp(left_rows), _vp(r_ifirstgroup), _vp(r_igroup),
                                   _vp(lo), _vp(ro), rc._stream()))
    return lo, ro


def _right_side(right: _KeyGroup, keep_right):
    """The right side as (nCountGroup, iFirstGroup, iGroup) indexed by 1-based key id.  With `keep` every key has exactly
    one row -- its first / last -- which makes the same expansion do `ifirstkey[keyid - 1]` (rt_merge.py:824-846)."""
    dev = right.dev
    if keep_right is None:
        cnt, ig, ifg = right.packed()
        return cnt, ifg, ig
    table = right.keep_rows(keep_right, False)
    u = right.u
    cnt = torch.ones(u + 1, dtype=torch.int32, device=dev)
    ifg = (torch.arange(u + 1, dtype=torch.int32, device=dev) - 1).clamp_(min=0)
    return cnt, ifg.contiguous(), (table if table.numel() else torch.zeros(4, dtype=torch.int32, device=dev))


def _left_or_inner(left: _KeyGroup, right: _KeyGroup, is_inner, keep, maps=None):
    """`left_or_inner_join` (rt_merge.py:1336-1402) = `_build_left_fancyindex` (:978-1190) + `_build_right_fancyindex`
    (:734-884), case by case."""
    keep_left, keep_right = keep
    dev = left.dev
    _, _, l2r1 = maps if maps is not None else _key_maps(left, right)
    nonmatched = 0 if is_inner else 1
    simple_right = keep_right is not None or right.u == right.n  # every key of right stands for exactly one row
    cnt, ifg, ig = _right_side(right, keep_right)

    # ---- right index: driven by all left rows, or by the kept row of every left group (the invalid group's included)
    rows_null = left.keep_rows(keep_left, True) if keep_left is not None else None
    ikey_r = left.ikey if keep_left is None else (left.ikey[rows_null.to(torch.int64)] if rows_null.numel() else _i32(0, dev))
    if ikey_r.numel() == 0:
        ri = _i32(0, dev)                                                   # rt_merge.py:791-797
    elif right.n == 0:
        ri = _i32(0, dev) if is_inner else torch.full((int(ikey_r.numel()),), INV, dtype=torch.int32, device=dev)  # :800-810
    else:
        _, ri = _expand(ikey_r, None, l2r1, cnt, ifg, ig, nonmatched, False, True)

    # ---- left index
    if simple_right and not is_inner:
        li = None if keep_left is None else rows_null                       # rt_merge.py:1083-1090
    elif keep_left is None:
        li, _ = _expand(left.ikey, None, l2r1, cnt, ifg, ig, nonmatched, True, False)      # :1068-1072, 1155-1175
    else:
        rows = left.keep_rows(keep_left, False)                             # :1076-1078, 1128, 1176-1180
        lk = left.ikey[rows.to(torch.int64)] if rows.numel() else _i32(0, dev)
        li, _ = _expand(lk, rows.contiguous(), l2r1, cnt, ifg, ig, nonmatched, True, False)
    return li, ri, None


def _rows_of_flagged_groups(g: _KeyGroup, flags):
    """Row numbers (ascending) of the rows whose group is flagged (flags: bool[U + 1] indexed by iKey, slot 0 = the invalid
    rows) == sorted(Grouping.extract_groups(flags, iGroup, ...)) (rt_merge.py:1452-1458, 1500-1503): the expansion with
    per-group counts 0 / 1 and the row itself as the payload."""
    dev = g.dev
    cnt = flags.to(torch.int32).contiguous()
    ident = torch.arange(1, cnt.numel(), dtype=torch.int32, device=dev)  # key id k -> "partner" k
    rows, _ = _expand(g.ikey, None, ident, cnt, cnt, cnt, 0, True, False, invalid_rows=int(cnt[0].item()))
    return rows


def _outer(left: _KeyGroup, right: _KeyGroup, right_gb, keep):
    """rt_merge.py:1404-1588: the left join, then the rows of right whose key does not occur in left (its invalid rows
    included), in right row order."""
    dev = left.dev
    maps = _key_maps(left, right)
    exists = maps[0]
    cnt, ig, _ = right.packed()
    n0 = int(cnt[0].item())
    only = torch.empty(right.u + 1, dtype=torch.bool, device=dev)
    only[0] = n0 > 0
    only[1:] = ~exists
    if keep[1] is None:
        rows_only = _rows_of_flagged_groups(right, only)
    else:
        if right_gb is not None and n0 > 0:
            inv_rows = ig[:n0].to(torch.int64)
            gb_groups = torch.unique(right_gb.ikey[inv_rows].to(torch.int64))
            table = right_gb.keep_rows(keep[1], True)
            has_null = int(right_gb.packed()[0][0].item()) > 0
            inv_picked = table[gb_groups if has_null else gb_groups - 1]
            valid_picked = right.keep_rows(keep[1], False)[only[1:]]
            rows_only = torch.cat([inv_picked, valid_picked])
        else:
            mask = only if n0 > 0 else only[1:]
            rows_only = right.keep_rows(keep[1], True)[mask]
        rows_only = torch.sort(rows_only).values.to(torch.int32)
    li, ri, _ = _left_or_inner(left, right, False, keep, maps)
    k = int(rows_only.numel())
    if k:
        if li is None:
            li = torch.arange(left.n, dtype=torch.int32, device=dev)
        li = torch.cat([li, torch.full((k,), INV, dtype=torch.int32, device=dev)])
        if ri is None:
            ri = torch.arange(right.n, dtype=torch.int32, device=dev)
        ri = torch.cat([ri, rows_only])
    return li, ri, k


def merge_indices(left_keys, right_keys, how: str = "left", keep=None):
    """rt.merge_indices on lists of key columns -> (left_index | None, right_index | None, right_only_rowcount | None)."""
    if isinstance(keep, str) or keep is None:
        keep = (keep, keep)
    keep = tuple(keep)
    for k in keep:
        if k not in (None, "first", "last"):
            raise ValueError("Invalid argument value passed for the `keep` parameter.")
    if how not in ("left", "right", "inner", "outer"):
        raise ValueError(f"Unsupported value ({how}) specified for the 'how' parameter.")
    if how == "outer" and keep[0] is not None:
        raise ValueError("Specifying a 'keep' value for the left Dataset (or both) with an outer merge is temporarily not "
                         "permitted due to a bug in the implementation.")  # riptable/rt_merge.py:2166-2170
    lcols_in, rcols_in = rc._as_list(left_keys), rc._as_list(right_keys)
    if len(lcols_in) != len(rcols_in) or not lcols_in:
        raise ValueError("merge: the two sides need the same, non-zero, number of key columns")
    device_mode = any(rc._is_device(c) for c in list(lcols_in) + list(rcols_in))
    lcols, rcols = [rc._to_dev(c) for c in lcols_in], [rc._to_dev(c) for c in rcols_in]
    for a, b in zip(lcols, rcols):
        if a.dtype != b.dtype and not (a.dtype.kind == b.dtype.kind and a.dtype.kind in "SU"):
            raise TypeError(f"merge: key columns must share a dtype ({a.dtype} vs {b.dtype}); riptable casts before the join")
    left, _ = _keygroups(lcols, how != "left", False)
    right, right_gb = _keygroups(rcols, how != "right", how == "outer")
    if how == "left":
        li, ri, k = _left_or_inner(left, right, False, keep)
    elif how == "right":
        ri, li, k = _left_or_inner(right, left, False, (keep[1], keep[0]))
        k = None
    elif how == "inner":
        li, ri, k = _left_or_inner(left, right, True, keep)
    elif left.n == 0 or right.n == 0:  # rt_merge.py:1600-1619
        if right.n == 0:
            li, ri, _ = _left_or_inner(left, right, False, keep)
            k = 0
        else:
            ri, li, _ = _left_or_inner(right, left, False, (keep[1], keep[0]))
            k = 0 if ri is None else int(ri.numel())
    else:
        li, ri, k = _outer(left, right, right_gb, keep)
    if not device_mode:
        li = None if li is None else li.cpu().numpy()
        ri = None if ri is None else ri.cpu().numpy()
    return li, ri, k
