"""The Python glue riptable wraps around the rc entry points (restated from riptable/rt_numpy.py:1901-1975,
1979-2094, 2098-2251), parameterised by the `rc` module so the same call sequence runs over the CPU oracle
and over riptable_b200.rc.  Test infrastructure only."""
import numpy as np


def groupbypack(rc, ikey, ncountgroup, unique_count=None, cutoffs=None):
    # rt_numpy.py:1945-1972
    if cutoffs is None:
        if ncountgroup is None:
            if unique_count is None or not np.isscalar(unique_count):
                raise ValueError("groupbypack: unique_count must be a scalar value if ncountgroup is None")
            if len(ikey) > 1e6 and len(ikey) / (unique_count + 1) < 40:
                ncountgroup = rc.BinCount(ikey, unique_count)
                igroup, ifirstgroup = rc.GroupFromBinCount(ikey, ncountgroup)
            else:
                ncountgroup, igroup, ifirstgroup = rc.BinCount(ikey, unique_count, pack=True)
        else:
            igroup, ifirstgroup = rc.GroupFromBinCount(ikey, ncountgroup)
    else:
        igroup, ifirstgroup, ncountgroup = rc.GroupByPack32(ikey, None, unique_count, cutoffs=cutoffs)
    return {"iGroup": igroup, "iFirstGroup": ifirstgroup, "nCountGroup": ncountgroup}


def groupbyhash(rc, list_arrays, hint_size=0, filter=None, hash_mode=2, cutoffs=None, pack=False):
    # rt_numpy.py:2064-2094
    if isinstance(list_arrays, np.ndarray):
        list_arrays = [list_arrays]
    if not (isinstance(list_arrays, list) and len(list_arrays) > 0):
        raise ValueError("groupbyhash first argument is not a list of numpy arrays")
    common_len = {len(arr) for arr in list_arrays}
    if len(common_len) != 1:
        raise ValueError(f"groupbyhash all arrays must have same length not {common_len}")
    if common_len.pop() != 0:
        ikey, ifirstkey, unique_count = rc.MultiKeyGroupBy32(list_arrays, hint_size, filter, hash_mode, cutoffs=cutoffs)
    else:
        ikey = np.zeros(0, np.int32)
        ifirstkey, unique_count = ikey, 0
    d = {"iKey": ikey, "iFirstKey": ifirstkey, "unique_count": unique_count}
    if pack:
        d.update(groupbypack(rc, ikey, None, unique_count + 1))
    else:
        d.update({"iGroup": None, "iFirstGroup": None, "nCountGroup": None})
    return d


def groupbylex(rc, list_arrays, filter=None, cutoffs=None, base_index=1, rec=False):
    # rt_numpy.py:2157-2251
    if base_index != 1 and base_index != 0:
        raise ValueError(f"Invalid base_index {base_index!r}")
    if base_index == 0 and filter is not None:
        raise ValueError("Filter and base_index of 0 cannot be combined")
    if isinstance(list_arrays, np.ndarray):
        list_arrays = [list_arrays]
    if len(list_arrays) > 1:
        value_array = np.rec.fromarrays(list_arrays)
        list_arrays = value_array if rec else list_arrays[::-1]
    else:
        value_array = list_arrays[0]
    index = None
    if filter is not None:
        filter = np.asarray(filter, dtype=bool)
        index = np.flatnonzero(filter).astype(np.int32)
        filterfalse = np.flatnonzero(~filter).astype(np.int32)
    iGroup = rc.LexSort32(list_arrays, cutoffs=cutoffs, index=index)
    retval = rc.GroupFromLexSort(iGroup, value_array, cutoffs=cutoffs, base_index=base_index)
    nUniqueCutoffs = None
    if len(retval) == 3:  # rt_numpy.py:2209-2212
        iKey, iFirstKey, nCountGroup = retval
    else:
        iKey, iFirstKey, nCountGroup, nUniqueCutoffs = retval
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
    d = {"iKey": iKey, "iFirstKey": iFirstKey, "unique_count": len(iFirstKey), "iGroup": iGroup,
         "iFirstGroup": iFirstGroup, "nCountGroup": nCountGroup}
    if nUniqueCutoffs is not None:
        d["nUniqueCutoffs"] = nUniqueCutoffs
    return d


def hstack_reindex(mod_groupby, mod_reindex, keys, cuts):
    """The hstack_groupings recipe (rt_grouping.py:277-413): per-partition grouping -> group the stacked uniques ->
    ReIndexGroups.  Returns (new ikey, global unique count)."""
    ik, (ifk, ucuts), _ = mod_groupby([keys], 0, None, 2, cutoffs=cuts)
    offs = np.repeat(np.r_[0, cuts[:-1]], np.diff(np.r_[0, ucuts]))
    stacked = np.asarray(keys)[np.asarray(ifk, np.int64) + offs]  # iFirstKey is partition-relative
    uik, _, U = mod_groupby([stacked], 0, None, 2)
    return mod_reindex(ik, uik, ucuts, cuts), U
