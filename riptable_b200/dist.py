"""Row-sharded grouping and reductions: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch).

This is riptable's own partition scheme, moved across GPUs.  riptable groups contiguous row partitions
independently (`cutoffs`, riptable/rt_numpy.py:1999-2000, riptable/tests/test_pgroupby.py:7-86) and then merges
the per-partition groupings by re-grouping the stacked unique keys (`hstack_groupings`,
riptable/rt_grouping.py:277-413: `groupbyhash(stacked uniques)` at :359-364, then `rc.ReIndexGroups` at :378-407).
Here rank r holds rows [offset_r, offset_r + n_r) and

  1. groups its shard on its own GPU                                  (rc.MultiKeyGroupBy32, a1)
  2. all-gathers every shard's unique key rows, in rank order         (ncclAllGather; U_r x key bytes per rank)
  3. groups the stacked uniques — every rank runs the same call       (rc.MultiKeyGroupBy32 again)
     Shards are in row order and local uniques are in first-appearance order, so first appearance in the
     stacked list IS global first appearance: the global bin ids and iFirstKey are bit-exact with a single-GPU run.
  4. re-indexes its local iKey through the local->global bin map      (rtb_remap_ikey, rc.ReIndexGroups semantics)

Reductions are associative, so each rank reduces its shard against the global bin ids and the partials are
combined: a plain all-reduce while the group count is small, and above `alltoall_min_bins` a bin-partitioned
all-to-all (rank r owns a contiguous slice of the bins), a local merge and an all-gather of the merged slices.
No data-path collective moves rows; only U-sized tables cross NVLink.

All device work goes through an `ops` object.  `CudaOps` (the product) drives riptable_b200.rc in device mode;
the gloo tests in tests/ supply an oracle-backed stand-in so the host logic and the collectives are covered
on CPU with world_size 2.  There is no CPU fallback inside this module.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as tdist

from .enums import GB_FUNCTIONS as GB

ALLTOALL_MIN_BINS = 8_000_000  # above this many bins the partials are merged by bin-partitioned all-to-all
SEEDED = os.environ.get("RTB_DIST_SEEDED", "1") == "1"
SEED_ROWS = 1 << 25  # rows of the first shard that are grouped before the other shards start
HOST_CHUNK_ROWS = 1 << 25  # HostFeed: rows per host <-> device chunk
SEEDED_MIN_UNIQUE = 1 << 20  # the seeded path provisions max(this, largest shard / 16) bins; beyond that it falls back


# ====================================================================================================
class CudaOps:
    """The product backend: riptable_b200.rc in device mode on this rank's GPU."""

    def __init__(self, device=None):
        from . import rc
        from ._lib import RtbWorkspaceError

        self.rc = rc
        self.workspace_error = (RtbWorkspaceError,)  # what a grouping that ran out of provisioned bins raises
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._side = None

    # -- columns ---------------------------------------------------------------------------------
    def as_col(self, x):
        return self.rc._to_dev(x)

    def nrows(self, col):
        return col.n

    def take_rows(self, col, idx):
        """col[idx] for any fixed-width dtype (a device gather of itemsize-byte rows)."""
        rc = self.rc
        isz = col.dtype.itemsize
        src = col.buf[: col.n * isz].view(col.n, isz)
        out = src.index_select(0, idx.to(torch.int64))
        k = int(idx.numel())
        buf = out.reshape(-1)
        if buf.numel() < 16:
            pad = torch.zeros(16, dtype=torch.uint8, device=buf.device)
            pad[: buf.numel()] = buf
            buf = pad
        return rc.DevArray(buf.contiguous(), col.dtype, k)

    def col_bytes(self, col):
        return col.buf[: col.n * col.dtype.itemsize]

    def col_from_bytes(self, t, dtype, n):
        rc = self.rc
        if t.numel() < 16:
            pad = torch.zeros(16, dtype=torch.uint8, device=t.device)
            pad[: t.numel()] = t
            t = pad
        return rc.DevArray(t.contiguous(), np.dtype(dtype), n)

    def col_dtype(self, col):
        return col.dtype

    # -- hot-path calls ----------------------------------------------------------------------------
    def groupby(self, cols, filter=None):
        ik, ifk, u = self.rc.MultiKeyGroupBy32([self._dev_view(c) for c in cols], 0, filter, 2)
        return ik, ifk, int(u)

    def _dev_view(self, c):
        return c  # DevArray is accepted by rc directly

    def remap(self, ikey, table):
        """out[i] = table[ikey[i]] on a side stream: the gather is bound by the L1TEX wavefront rate while the
        reductions that follow are bound by the L2 atomic unit, so the two overlap well.  Returns (out, event)."""
        from ._lib import lib, check
        import ctypes as C

        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream()
        out = torch.empty_like(ikey)
        self._side.wait_stream(cur)
        for t in (ikey, table, out):
            t.record_stream(self._side)
        check(lib.rtb_remap_ikey(C.c_void_p(ikey.data_ptr()), ikey.numel(), C.c_void_p(table.data_ptr()), table.numel(),
                                 C.c_void_p(out.data_ptr()), C.c_void_p(self._side.cuda_stream)))
        ev = torch.cuda.Event()
        ev.record(self._side)
        return out, ev

    def wait(self, ev):
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def stream(self, max_unique, max_chunk_rows, with_sums=False):
        """Resumable single-key grouping (rc.GroupByStream): push(keys, filter, row_offset, out[, values]) -> iKey of
        the chunk; with_sums: the pushes also add the chunk's values to per-bin accumulators (stream.sums())."""
        return self.rc.GroupByStream(max_unique, max_chunk_rows, with_sums=with_sums)

    def can_stream(self, col):
        return col.dtype.kind in "biuf" and col.dtype.itemsize in (1, 2, 4, 8)

    def slice_rows(self, col, lo, hi):
        isz = col.dtype.itemsize
        return self.rc.DevArray(col.buf[lo * isz:], col.dtype, hi - lo) if hi > lo else self.rc.DevArray.empty(0, col.dtype)

    # -- adopting a seed (dist._sharded_groupby_seeded): U-sized tables, plus one re-index of the prefix rows -----------
    def bin_table(self, newbins):
        """[0, new bin of old bin 1, new bin of old bin 2, ...] as int32."""
        t = torch.empty(newbins.numel() + 1, dtype=torch.int32, device=newbins.device)
        t[0] = 0
        t[1:] = newbins
        return t

    def stream_fix_first(self, st, u0, u_mid, old_first):
        """First rows of bins (u0, u_mid] were recorded as positions in the chunk of local unique keys: turn them into rows."""
        f = st.ifirst.torch()
        f[u0:u_mid] = old_first[f[u0:u_mid].to(torch.int64)].to(torch.int32)

    def remap_prefix(self, ikey, S, table):
        """ikey[:S] <- table[ikey[:S]] (rtb_remap_ikey); the rest of the buffer is still to be written."""
        if S == 0:
            return ikey
        from ._lib import lib, check
        import ctypes as C

        # in place: every thread reads its four rows before it writes them
        check(lib.rtb_remap_ikey(C.c_void_p(ikey.data_ptr()), int(S), C.c_void_p(table.data_ptr()), table.numel(),
                                 C.c_void_p(ikey.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return ikey

    def stream_move_sums(self, old, new, table, u_loc):
        """Partial sums of the prefix, accumulated under the local bins, move to the adopted bins (a permutation)."""
        src = old._acc[: u_loc + 1]
        new._vdtype = old._vdtype
        new._acc[table.to(torch.int64)] = src

    def slice_filter(self, f, lo, hi):
        return None if f is None else f[lo:hi]

    def empty_ikey(self, n):
        return torch.empty(max(n, 4), dtype=torch.int32, device=self.device)

    def reduce(self, val, ikey, nuniq, func):
        return self.rc.GroupByAll32([val], ikey, nuniq, [int(func)], [0], [nuniq + 1], 0)[0]

    def bincount(self, ikey, nbins):
        return self.rc.BinCount(ikey, nbins)

    def torch_dtype(self, col):
        return col.torch().dtype

    def valid_mask_as_i8(self, col):
        t = col.torch()
        return (~_invalid_mask(t)).to(torch.int8)

    def ifirst(self, ikey, nuniq, mode):
        return self.rc.MakeiFirst(ikey, nuniq, None, mode)

    def pack(self, ikey, nbins):
        return self.rc.BinCount(ikey, nbins, pack=True)  # (nCountGroup, iGroup, iFirstGroup)

    def cumsum(self, val, ikey, nuniq, filter):
        return self.rc.EmaAll32([val], ikey, nuniq, int(GB.GB_CUMSUM), (0.0, None, filter, None))[0]

    def combine_filter(self, ikey, filter):
        return self.rc.CombineFilter(ikey, filter)

    def as_f64(self, col):
        return col.torch().to(torch.float64)

    def lexsort(self, cols, ascending=True):
        return self.rc.LexSort32(list(cols), ascending=ascending)

    def lex_partition(self, cols, splitters, splitter_rows, row_offset, ascending):
        return self.rc.LexPartition(list(cols), list(splitters), splitter_rows, row_offset, ascending)

    def group_from_sorted(self, cols):
        """rc.GroupFromLexSort over columns that are already in sorted order -> (bin per position, base 1;
        first position of each bin; nCountGroup)."""
        ident = torch.arange(cols[0].n, dtype=torch.int32, device=self.device)
        return self.rc.GroupFromLexSort(ident, list(cols))

    def rows_col(self, rows_i64):
        """An int64 tensor of global row ids as a key column."""
        return self.rc._to_dev(rows_i64.contiguous())

    def is_f32(self, col):
        return col.dtype == np.dtype(np.float32)

    def invalid_item_bytes(self, col):
        """The dtype's invalid as raw bytes (riptable/rt_enum.py:88-143): NaN, integer sentinel, b'' / ''."""
        from .enums import invalid_for
        dt = col.dtype
        if dt.kind in "SUV":
            return torch.zeros(dt.itemsize, dtype=torch.uint8, device=self.device)
        return torch.from_numpy(np.array([invalid_for(dt)], dtype=dt).view(np.uint8).copy()).to(self.device)

    def tensor(self, x, dtype=None):
        return torch.as_tensor(x, dtype=dtype, device=self.device)

    def to_torch(self, x):
        return x  # device-mode rc results are torch tensors already


# ====================================================================================================
def _world(group):
    return tdist.get_world_size(group), tdist.get_rank(group)


def _all_gather_var(t: torch.Tensor, counts, group):
    """All-gather 1-D tensors of different lengths (counts[r] elements from rank r), concatenated in rank order."""
    G = len(counts)
    m = int(max(counts)) if G else 0
    if m == 0:
        return t.new_empty(0)
    pad = t.new_zeros(m)
    pad[: t.numel()] = t
    out = t.new_empty(G * m)
    tdist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m: r * m + int(counts[r])] for r in range(G)])


class ShardedGrouping:
    """Result of `sharded_groupby`: global bin ids for this rank's rows plus the replicated group table."""

    def __init__(self, ops, gikey, gikey_ready, local_ikey, local_unique_count, table, ifirstkey, unique_count, row_offset,
                 nrows_total, shard_unique_counts):
        self._ops, self._gikey, self._ready = ops, gikey, gikey_ready
        self.local_ikey = local_ikey        # int32[n_r]  this shard's own first-appearance bins (what reductions read)
        self.local_unique_count = local_unique_count
        self.table = table                  # int32[u_r + 1] local bin -> global bin (table[0] = 0)
        self.ifirstkey = ifirstkey          # int64[U]    GLOBAL row of each bin's first appearance (replicated)
        self.unique_count = unique_count    # U
        self.row_offset = row_offset        # first global row of this shard
        self.nrows_total = nrows_total
        self.shard_unique_counts = shard_unique_counts
        self.partial_sums = None            # raw per-bin sums of this shard by GLOBAL bin, when the grouping pass made them
        self.host_ikey_streamed = False     # a HostFeed already holds the final iKey of every chunk

    @property
    def ikey(self):
        """int32[n_r] global base-1 bin of every local row (0 = filtered); waits for the re-index pass."""
        if self._ready is not None:
            self._ops.wait(self._ready)
            self._ready = None
        return self._gikey


def sharded_groupby(ops, key_cols, filter=None, group=None, sum_values=None, shard_sizes=None, feed=None) -> ShardedGrouping:
    """MultiKeyGroupBy32 over rows sharded contiguously across the ranks of `group` (module docstring, steps 1-4).
    `sum_values`: a value column to sum in the same pass where the path allows it (single-key seeded shards); the
    grouping then carries `partial_sums` (this shard's raw per-bin sums by GLOBAL bin) for sharded_groupby_sum.
    `shard_sizes`: the row count of every rank's shard, when the caller knows them (saves a collective and a host sync)."""
    G, rank = _world(group)
    cols = [ops.as_col(c) for c in key_cols]
    n = ops.nrows(cols[0])
    if len(cols) == 1 and SEEDED and hasattr(ops, "stream") and ops.can_stream(cols[0]):
        g = _sharded_groupby_seeded(ops, cols[0], filter, group, sum_values, shard_sizes, feed)
        if g is not None:
            return g
        # a shard holds more unique keys than the seeded table was provisioned for: every rank takes the generic path
    if feed is not None:
        feed.wait_rows(n)  # the generic path reads the whole shard
    tick = _Tick()
    ik, ifk, u = ops.groupby(cols, filter)
    ik, ifk = ops.to_torch(ik), ops.to_torch(ifk)
    tick.mark("local groupby")

    # shard sizes and unique counts of every rank
    meta = ops.tensor([n, u], dtype=torch.int64)
    allmeta = meta.new_empty(2 * G)
    tdist.all_gather_into_tensor(allmeta, meta, group=group)
    allmeta = allmeta.cpu().view(G, 2)
    ns, us = allmeta[:, 0].tolist(), allmeta[:, 1].tolist()
    row_off = [0] * G
    for r in range(1, G):
        row_off[r] = row_off[r - 1] + ns[r - 1]
    uoff = [0] * G
    for r in range(1, G):
        uoff[r] = uoff[r - 1] + us[r - 1]
    total_u = uoff[-1] + us[-1]

    # stacked unique key rows (rank order == global row order) and their global first rows
    stacked = []
    for c in cols:
        dt = ops.col_dtype(c)
        uniq = ops.take_rows(c, ifk[:u])
        allb = _all_gather_var(ops.col_bytes(uniq), [x * dt.itemsize for x in us], group)
        stacked.append(ops.col_from_bytes(allb, dt, total_u))
    gfirst_local = ifk[:u].to(torch.int64) + row_off[rank]
    all_first = _all_gather_var(gfirst_local, us, group)
    tick.mark("gather uniques")

    if total_u == 0:
        empty64 = all_first.new_empty(0)
        return ShardedGrouping(ops, ik, None, ik, 0, torch.zeros(1, dtype=torch.int32, device=ik.device), empty64, 0,
                               row_off[rank], sum(ns), us)

    gk, gfk, U = ops.groupby(stacked, None)
    gk, gfk = ops.to_torch(gk), ops.to_torch(gfk)
    tick.mark("group stacked uniques")
    # local bin b (1..u) -> global bin of stacked row uoff[rank] + b - 1; bin 0 stays 0
    table = torch.zeros(u + 1, dtype=torch.int32, device=gk.device)
    table[1:] = gk[uoff[rank]: uoff[rank] + u]
    if rank == 0:
        gikey, ready = ik, None  # the first shard's uniques lead the stacked list: its local bins are the global bins
    else:
        gikey, ready = ops.remap(ik, table)
    tick.mark("remap ikey (enqueue)")
    ifirst_global = all_first.index_select(0, gfk[:U].to(torch.int64))
    tick.mark("ifirstkey")
    tick.report()
    return ShardedGrouping(ops, gikey, ready, ik, u, table, ifirst_global, int(U), row_off[rank], sum(ns), us)


def _sharded_groupby_seeded(ops, col, filter, group, sum_values=None, shard_sizes=None, feed=None) -> ShardedGrouping:
    """Single-key shards, seeded: global bin ids come out of the hash pass itself, no per-row re-index.

    Global first-appearance order lists every key of shard 0 first (in shard 0's own order).  Shard 0 therefore groups
    a prefix of its rows and broadcasts the keys it found (u0 of them, typically all) with their bins 1..u0; every other
    rank hashes its rows against a table that holds exactly those keys under exactly those bins (the resumable grouping
    of rc.GroupByStream), so rows whose key is in the seed get their GLOBAL bin straight out of the hash pass.

    The other ranks do not wait for the seed: while shard 0 groups its prefix they group a prefix of their own under
    local bins, and when the seed arrives they ADOPT it -- a fresh table is loaded with the seed keys, the local unique
    keys are pushed through it (one U-sized chunk: seeded keys come back with their global bin, the others get bins above
    u0 in local first-appearance order), and the prefix's iKey and partial sums are re-indexed through that U-sized map.
    Only the prefix is re-indexed; the rest of the shard is hashed against the adopted table.

    `feed` (HostFeed): the shard is still arriving from host memory.  Rows are pushed chunk by chunk as they land, and
    every chunk's iKey is handed back as soon as it is final (the prefix after the adoption, the rest after its push).

    Keys outside the seed ("late" keys: first seen after the prefix of shard 0, or absent from shard 0) end up with
    provisional bins above u0; only if any rank reports such keys are they merged (all-gather of the late uniques in
    rank order, re-group, re-index)."""
    G, rank = _world(group)
    n = ops.nrows(col)
    dt = ops.col_dtype(col)
    tick = _Tick()
    if shard_sizes is not None:
        ns = [int(x) for x in shard_sizes]  # the caller knows the shard sizes: no collective, no host sync
    else:
        meta = ops.tensor([n], dtype=torch.int64)
        alln = meta.new_empty(G)
        tdist.all_gather_into_tensor(alln, meta, group=group)
        ns = alln.cpu().tolist()
    row_off = [0] * G
    for r in range(1, G):
        row_off[r] = row_off[r - 1] + ns[r - 1]
    nmax = max(ns) if ns else 0
    max_unique = max(SEEDED_MIN_UNIQUE, nmax // 16, 1)
    S = (min(n, SEED_ROWS) & ~127) if n > SEED_ROWS else n  # this rank's prefix
    fused = sum_values is not None
    vcol = ops.as_col(sum_values) if fused else None

    def new_stream():
        return ops.stream(max_unique, max(nmax, 1), with_sums=True) if fused else ops.stream(max_unique, max(nmax, 1))

    st = new_stream()
    ikey = ops.empty_ikey(n)

    def push(s_, lo_, hi_, out):
        if fused:
            s_.push(ops.slice_rows(col, lo_, hi_), ops.slice_filter(filter, lo_, hi_), lo_, out, values=ops.slice_rows(vcol, lo_, hi_))
        else:
            s_.push(ops.slice_rows(col, lo_, hi_), ops.slice_filter(filter, lo_, hi_), lo_, out)

    # A push that runs out of table room fails on that rank alone; the failure is carried by the next collective
    # every rank takes part in (u0 = -1 / late = -1), so that all ranks leave the seeded path together.
    overflow = getattr(ops, "workspace_error", ())

    # 1. every rank groups its own prefix; shard 0's is the seed
    failed = False
    try:
        if S:
            if feed is not None:
                feed.wait_rows(S)
            push(st, 0, S, ikey[:S])
    except overflow:
        failed = True
    tick.mark("own prefix")
    src = tdist.get_global_rank(group, 0) if group is not None else 0
    u0_t = ops.tensor([(-1 if failed else st.unique_count) if rank == 0 else 0], dtype=torch.int64)
    tdist.broadcast(u0_t, src=src, group=group)
    u0 = int(u0_t.item())
    tick.mark("u0 broadcast")
    if u0 < 0:
        return None
    # one payload: the seed's key bytes followed by its first rows
    isz = dt.itemsize
    payload = ops.tensor([0], dtype=torch.uint8).new_empty(u0 * (isz + 8))
    if rank == 0 and u0:
        first0 = ops.to_torch(st.ifirstkey).to(torch.int64)
        payload[: u0 * isz] = ops.col_bytes(ops.take_rows(col, first0))
        payload[u0 * isz:] = first0.contiguous().view(torch.uint8)
    if u0:
        tdist.broadcast(payload, src=src, group=group)
    seed_bytes = payload[: u0 * isz]
    seed_first = payload[u0 * isz:].clone().view(torch.int64) if u0 else ops.tensor([0], dtype=torch.int64).new_empty(0)
    tick.mark("seed broadcast")

    # 2. the other ranks adopt the seed, then all ranks hash the rest of their rows
    try:
        if rank != 0 and not failed:
            seed_col = ops.col_from_bytes(seed_bytes, dt, u0)
            table = st.adopt_seed(seed_col) if (u0 and hasattr(st, "adopt_seed")) else None
            if table is not None:
                # the table was re-labelled in place (first rows and partial sums with it): only the prefix's iKey is left
                ikey = ops.remap_prefix(ikey, S, ops.to_torch(table))
            else:
                # no room in the local table (or no local prefix): a fresh table loaded with the seed, the local unique
                # keys pushed through it as one chunk, and the prefix re-indexed through the resulting U-sized map
                st2 = new_stream()
                if u0:
                    st2.push(seed_col, None, 0, None)
                    if st2.unique_count != u0:
                        raise RuntimeError("sharded_groupby: the broadcast seed holds duplicate keys")
                u_loc = st.unique_count
                if u_loc:
                    old_first = ops.to_torch(st.ifirstkey).to(torch.int64).clone()
                    newbins = st2.push(ops.take_rows(col, old_first), None, 0, None)   # local unique keys, local order
                    table = ops.bin_table(newbins)                                     # [0, new bin of local bin 1, ...]
                    u_mid = st2.unique_count
                    if u_mid > u0:  # first rows of the local keys the seed does not hold: positions in the chunk -> rows
                        ops.stream_fix_first(st2, u0, u_mid, old_first)
                    ikey = ops.remap_prefix(ikey, S, table)
                    if fused:
                        ops.stream_move_sums(st, st2, table, u_loc)
                st = st2
        tick.mark("adopt seed")
        if feed is not None and S and not failed:
            feed.ikey_final(0, S, ikey)
        if n > S and not failed:
            if feed is None:
                push(st, S, n, ikey[S:n])
            else:
                for lo_, hi_ in feed.chunks(S, n):
                    feed.wait_rows(hi_)
                    push(st, lo_, hi_, ikey[lo_:hi_])
                    feed.ikey_final(lo_, hi_, ikey)
    except overflow:
        failed = True
    ikey = ikey[:n]
    U_r = st.unique_count
    tick.mark("hash shard")

    # 3. late keys (rare): merge them across ranks
    late = ops.tensor([-1 if failed else U_r - u0], dtype=torch.int64)
    lates = late.new_empty(G)
    tdist.all_gather_into_tensor(lates, late, group=group)
    lates = lates.cpu().tolist()
    if min(lates) < 0:
        return None
    total_late = sum(lates)
    if total_late == 0:
        tick.mark("late check")
        tick.report()
        g = ShardedGrouping(ops, ikey, None, ikey, u0, None, seed_first, u0, row_off[rank], sum(ns), [u0] * G)
        if fused:
            g.partial_sums = ops.to_torch(st.sums())  # bins are global already: ready for the all-reduce
        g.host_ikey_streamed = feed is not None  # every chunk's iKey went to the host as it became final
        return g
    ifk = ops.to_torch(st.ifirstkey)
    late_rows = ifk[u0:U_r]
    late_keys = ops.take_rows(col, late_rows)
    stacked = ops.col_from_bytes(_all_gather_var(ops.col_bytes(late_keys), [x * dt.itemsize for x in lates], group), dt, total_late)
    all_first = _all_gather_var(late_rows.to(torch.int64) + row_off[rank], lates, group)
    gk, gfk, L = ops.groupby([stacked], None)
    gk, gfk = ops.to_torch(gk), ops.to_torch(gfk)
    loff = sum(lates[:rank])
    table = torch.arange(U_r + 1, dtype=torch.int32, device=gk.device)
    table[u0 + 1:] = gk[loff: loff + lates[rank]] + u0
    if rank == 0:
        gikey, ready = ikey, None  # shard 0's late keys lead the stacked list: already global
    else:
        gikey, ready = ops.remap(ikey, table)
    ifirst_global = torch.cat([seed_first, all_first.index_select(0, gfk[:L].to(torch.int64))])
    U = u0 + int(L)
    tick.mark("late merge")
    tick.report()
    g = ShardedGrouping(ops, gikey, ready, None, U, None, ifirst_global, U, row_off[rank], sum(ns), [u0 + x for x in lates])
    g.local_ikey = g.ikey  # reductions read the merged bins (this waits for the re-index pass)
    return g


# ----------------------------------------------------------------------------------------------------
def _invalid_mask(t: torch.Tensor):
    if t.dtype.is_floating_point:
        return torch.isnan(t)
    if t.dtype == torch.bool:
        return torch.zeros_like(t, dtype=torch.bool)
    info = torch.iinfo(t.dtype)
    sentinel = info.min if info.min < 0 else info.max
    return t == sentinel


def _invalid_value(dtype):
    if dtype.is_floating_point:
        return float("nan")
    if dtype == torch.bool:
        return False
    info = torch.iinfo(dtype)
    return info.min if info.min < 0 else info.max


def _merge_slices(kind, parts, nan_op):
    """Merge per-rank partials of one bin slice.  parts: dict name -> tensor [G, m].  Returns dict name -> [m]."""
    if kind == "sum":
        return {"v": parts["v"].sum(dim=0)}
    if kind == "count":
        return {"v": parts["v"].sum(dim=0, dtype=torch.int64).to(torch.int32)}
    if kind == "mean":
        c = parts["c"].sum(dim=0)
        s = parts["s"].sum(dim=0)
        return {"v": torch.where(c > 0, s / c.to(torch.float64), torch.full_like(s, float("nan")))}
    if kind == "var":
        # Chan et al. pairwise merge of (count, mean, M2), folded over the ranks in rank order
        c, mean, m2 = parts["c"][0].to(torch.float64), parts["mean"][0].clone(), parts["m2"][0].clone()
        mean = torch.where(c > 0, mean, torch.zeros_like(mean))
        for r in range(1, parts["c"].shape[0]):
            cb = parts["c"][r].to(torch.float64)
            mb = torch.where(cb > 0, parts["mean"][r], torch.zeros_like(mean))
            tot = c + cb
            delta = mb - mean
            safe = torch.where(tot > 0, tot, torch.ones_like(tot))
            mean = mean + delta * cb / safe
            m2 = m2 + parts["m2"][r] + delta * delta * c * cb / safe
            c = tot
        return {"c": c, "m2": m2}
    if kind in ("min", "max"):
        a, cnt = parts["a"], parts["c"]
        inv = _invalid_mask(a)
        seen_invalid = (inv & (cnt > 0)).any(dim=0) if not nan_op else None
        if a.dtype.is_floating_point:
            fill = float("inf") if kind == "min" else float("-inf")
        elif a.dtype == torch.bool:
            a = a.to(torch.uint8)
            fill = 1 if kind == "min" else 0
        else:
            info = torch.iinfo(a.dtype)
            fill = info.max if kind == "min" else info.min
        filled = torch.where(inv, torch.full_like(a, fill), a)
        red = filled.min(dim=0).values if kind == "min" else filled.max(dim=0).values
        any_valid = (~inv).any(dim=0)
        bad = ~any_valid if nan_op else (~any_valid | seen_invalid)
        out = torch.where(bad, torch.full_like(red, _invalid_value(parts["a"].dtype) if parts["a"].dtype != torch.bool else 0), red)
        return {"v": out.to(parts["a"].dtype)}
    raise ValueError(kind)


def _combine(kind, local: dict, nbins: int, nan_op: bool, group, alltoall_min_bins):
    """Combine per-rank [nbins] partials into the replicated merged result(s)."""
    G, rank = _world(group)
    names = sorted(local)
    if nbins < alltoall_min_bins and kind in ("sum", "count"):
        v = local["v"].clone()
        tdist.all_reduce(v, op=tdist.ReduceOp.SUM, group=group)
        return {"v": v}
    if nbins < alltoall_min_bins:
        # small tables: all-gather the partials and merge everything locally
        parts = {}
        for nm in names:
            t = local[nm].contiguous()
            out = t.new_empty((G,) + tuple(t.shape))
            tdist.all_gather_into_tensor(out.view(-1), t.view(-1), group=group)
            parts[nm] = out
        return _merge_slices(kind, parts, nan_op)
    # large tables: rank r owns bins [r*m, (r+1)*m); all-to-all the slices, merge, all-gather the merged slices
    m = (nbins + G - 1) // G
    parts = {}
    for nm in names:
        t = local[nm]
        pad = t.new_zeros(G * m)
        pad[:nbins] = t
        recv = torch.empty_like(pad)
        tdist.all_to_all_single(recv, pad, group=group)
        parts[nm] = recv.view(G, m)
    merged = _merge_slices(kind, parts, nan_op)
    out = {}
    for nm, t in merged.items():
        full = t.new_empty(G * m)
        tdist.all_gather_into_tensor(full, t.contiguous(), group=group)
        out[nm] = full[:nbins]
    return out


def sharded_reduce(ops, values, grouping: ShardedGrouping, func, group=None, alltoall_min_bins=ALLTOALL_MIN_BINS):
    """GroupByAll32 (sum, mean, min, max, var, std and the nan-variants) over the sharded rows; the merged
    [U+1] result is replicated on every rank.  Per-op semantics as riptable/rt_grouping.py:3433-3436."""
    func = int(func)
    U = grouping.unique_count
    nb = U + 1
    nan_op = func >= 50
    base = func - 50 if nan_op else func
    ik = grouping.local_ikey          # reduce against the shard's own bins; only U-sized partials are re-indexed
    lu = grouping.local_unique_count
    table = grouping.table.to(torch.int64) if grouping.table is not None else None  # None: local bins are global
    val = ops.as_col(values)

    def to_global(local, fill):
        if table is None:
            return local
        out = torch.full((nb,), fill, dtype=local.dtype, device=local.device)
        out[table] = local
        return out

    def red(f, fill):
        return to_global(ops.to_torch(ops.reduce(val, ik, lu, f)), fill)

    def counts(valid_only):
        return to_global(_valid_counts(ops, val, ik, lu, valid_only), 0)

    if base == GB.GB_SUM:
        return _combine("sum", {"v": red(func, 0)}, nb, nan_op, group, alltoall_min_bins)["v"]
    if base == GB.GB_MEAN:
        # per-bin (sum, count of accumulated rows): count = rows that a nan-op would keep
        s = red(GB.GB_NANSUM if nan_op else GB.GB_SUM, 0).to(torch.float64)
        c = counts(nan_op)
        return _combine("mean", {"s": s, "c": c}, nb, nan_op, group, alltoall_min_bins)["v"]
    if base in (GB.GB_MIN, GB.GB_MAX):
        odt = ops.torch_dtype(val)  # min/max keep the value dtype
        a = red(func, _invalid_value(odt))
        c = counts(False)
        kind = "min" if base == GB.GB_MIN else "max"
        return _combine(kind, {"a": a, "c": c}, nb, nan_op, group, alltoall_min_bins)["v"]
    if base in (GB.GB_VAR, GB.GB_STD):
        c = counts(nan_op)
        mean = red(GB.GB_NANMEAN if nan_op else GB.GB_MEAN, float("nan"))
        var = red(GB.GB_NANVAR if nan_op else GB.GB_VAR, float("nan"))  # ddof = 1 locally
        cf = c.to(torch.float64)
        m2 = torch.where(c >= 2, var * (cf - 1.0), torch.zeros_like(var))  # NaN var (a NaN input) stays NaN
        if not nan_op:
            poisoned = torch.isnan(mean) & (c > 0)
            m2 = torch.where(poisoned, torch.full_like(m2, float("nan")), m2)
        out = _combine("var", {"c": c, "mean": mean, "m2": m2}, nb, nan_op, group, alltoall_min_bins)
        v = torch.where(out["c"] >= 2, out["m2"] / (out["c"] - 1.0), torch.full_like(out["m2"], float("nan")))
        return torch.sqrt(v) if base == GB.GB_STD else v
    raise NotImplementedError(f"sharded_reduce: function {func}")


def _valid_counts(ops, val, ik, U, nan_op):
    """Rows per bin that the op accumulates: all rows of the bin, or only the valid ones for a nan-op."""
    if not nan_op:
        return ops.to_torch(ops.bincount(ik, U + 1))
    ones = ops.valid_mask_as_i8(val)  # 1 where the value is neither NaN nor the dtype's sentinel
    return ops.to_torch(ops.reduce(ones, ik, U, GB.GB_SUM)).to(torch.int32)


def sharded_count(ops, grouping: ShardedGrouping, group=None, alltoall_min_bins=ALLTOALL_MIN_BINS):
    """Grouping.count across shards: rc.BinCount per shard, summed."""
    c = ops.to_torch(ops.bincount(grouping.local_ikey, grouping.local_unique_count + 1))
    if grouping.table is None:
        full = c
    else:
        full = torch.zeros(grouping.unique_count + 1, dtype=c.dtype, device=c.device)
        full[grouping.table.to(torch.int64)] = c
    return _combine("count", {"v": full}, grouping.unique_count + 1, False, group, alltoall_min_bins)["v"]


_TRACE = os.environ.get("RTB_DIST_TRACE") == "1"  # per-phase device times of every sharded step, per rank


class _Tick:
    """CUDA-event stopwatch for RTB_DIST_TRACE=1 (prints the per-phase device time of one sharded step on rank 0)."""

    def __init__(self):
        self.ev, self.names = [], []
        self.mark("start")

    def mark(self, name):
        if _TRACE:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.ev.append(e)
            self.names.append(name)

    def report(self):
        if _TRACE:
            torch.cuda.synchronize()
            parts = [f"{self.names[i]} {self.ev[i - 1].elapsed_time(self.ev[i]):.2f}" for i in range(1, len(self.ev))]
            print(f"[rtb dist r{tdist.get_rank()}] ms: " + ", ".join(parts), flush=True)


def sharded_pick(ops, values, grouping: ShardedGrouping, func, nth=0, group=None):
    """GroupByAllPack32 first / nth / last over the sharded rows (riptable/rt_grouping.py:3442-3454): the replicated
    [U+1] result, invalid (NaN / sentinel / empty string) where a bin has no such row.  A bin's rows are ordered by
    global row = rank order, then local order."""
    func = int(func)
    G, rank = _world(group)
    U = grouping.unique_count
    nb = U + 1
    ik = grouping.ikey
    val = ops.as_col(values)
    dt = ops.col_dtype(val)
    isz = dt.itemsize
    if func in (GB.GB_FIRST, GB.GB_LAST):
        rows = ops.to_torch(ops.ifirst(ik, U, 0 if func == GB.GB_FIRST else 1)).to(torch.int64)  # [U+1], < 0 = none
    elif func == GB.GB_NTH:
        cnt, ig, ifg = (ops.to_torch(x) for x in ops.pack(ik, nb))
        cnts = cnt.new_empty(G * nb)
        tdist.all_gather_into_tensor(cnts, cnt.contiguous(), group=group)
        cnts = cnts.view(G, nb).to(torch.int64)
        before = cnts[:rank].sum(dim=0)
        total = cnts.sum(dim=0)
        want = torch.full_like(total, int(nth)) if nth >= 0 else total + int(nth)
        local = want - before  # index inside this rank's rows of the bin
        mine = (want >= 0) & (want < total) & (local >= 0) & (local < cnt.to(torch.int64))
        pos = (ifg.to(torch.int64) + local).clamp(0, max(int(ig.numel()) - 1, 0))
        rows = torch.where(mine, ig.to(torch.int64)[pos] if ig.numel() else torch.zeros_like(pos), torch.full_like(pos, -1))
    else:
        raise NotImplementedError(f"sharded_pick: function {func}")
    rows[0] = -1
    have = rows >= 0
    if ops.nrows(val) == 0:
        picked = ops.invalid_item_bytes(val).unsqueeze(0).expand(nb, isz).contiguous()  # an empty shard offers nothing
    else:
        picked = ops.col_bytes(ops.take_rows(val, rows.clamp(min=0))).view(nb, isz)
    big = torch.iinfo(torch.int64).max
    if func == GB.GB_LAST:
        key = torch.where(have, rows + grouping.row_offset, torch.full_like(rows, -1))
    else:
        key = torch.where(have, rows + grouping.row_offset, torch.full_like(rows, big))
    keys = key.new_empty(G * nb)
    tdist.all_gather_into_tensor(keys, key.contiguous(), group=group)
    vals = picked.new_empty(G * nb * isz)
    tdist.all_gather_into_tensor(vals, picked.contiguous().view(-1), group=group)
    keys, vals = keys.view(G, nb), vals.view(G, nb, isz)
    win = keys.argmax(dim=0) if func == GB.GB_LAST else keys.argmin(dim=0)
    best = keys.gather(0, win.unsqueeze(0)).squeeze(0)
    found = (best >= 0) if func == GB.GB_LAST else (best < big)
    out = vals.gather(0, win.view(1, nb, 1).expand(1, nb, isz)).squeeze(0)
    out = torch.where(found.unsqueeze(1), out, ops.invalid_item_bytes(val).unsqueeze(0).expand(nb, isz))
    return ops.col_from_bytes(out.contiguous().view(-1), dt, nb)


def sharded_cumsum(ops, values, grouping: ShardedGrouping, filter=None, group=None):
    """EmaAll32 GB_CUMSUM over the sharded rows (riptable/rt_grouping.py:3427-3430): every rank gets the running sums
    of its own rows.  Local running sums + the per-bin totals of all earlier shards (an exclusive scan over ranks of
    U-sized tables); op-filter rows carry the running value, bin-0 rows stay invalid."""
    G, rank = _world(group)
    U = grouping.unique_count
    nb = U + 1
    ik = grouping.ikey
    val = ops.as_col(values)
    f32 = ops.is_f32(val)
    work = ops.as_f64(val) if f32 else val  # float32 accumulates in float64, rounded once at the end
    local = ops.to_torch(ops.cumsum(work, ik, U, filter))
    ikf = ops.to_torch(ops.combine_filter(ik, filter)) if filter is not None else ik
    tot = ops.to_torch(ops.reduce(work, ikf, U, GB.GB_SUM)).to(local.dtype)
    tot[0] = 0
    tots = tot.new_empty(G * nb)
    tdist.all_gather_into_tensor(tots, tot.contiguous(), group=group)
    before = tots.view(G, nb)[:rank].sum(dim=0) if rank > 0 else torch.zeros_like(tot)
    out = local + before[ik.to(torch.int64)]
    if not local.dtype.is_floating_point:
        inv = _invalid_value(local.dtype)
        out = torch.where(ik == 0, torch.full_like(out, inv), out)
    return out.to(torch.float32) if f32 else out


def sharded_pack(ops, grouping: ShardedGrouping, group=None):
    """groupbypack over sharded rows (riptable/rt_numpy.py:1945-1972): the packed layout distributed by bin owner.
    Every rank gets the replicated global `nCountGroup` / `iFirstGroup` (int64: global positions), the bin range
    [bin_lo, bin_hi) it owns (contiguous ranges balanced by row count) and `iGroup_owned` = the global row ids of
    those bins in packed order (bin-major, global rows ascending inside a bin) — i.e. the slice
    iGroup[iFirstGroup[bin_lo] : iFirstGroup[bin_hi]] of the single-GPU result."""
    G, rank = _world(group)
    U = grouping.unique_count
    nb = U + 1
    ik = grouping.ikey
    cnt, ig, ifg = (ops.to_torch(x) for x in ops.pack(ik, nb))
    cnts = cnt.new_empty(G * nb)
    tdist.all_gather_into_tensor(cnts, cnt.contiguous(), group=group)
    cnts = cnts.view(G, nb).to(torch.int64)
    total = cnts.sum(dim=0)
    csum = torch.cumsum(total, 0)
    gifg = csum - total  # exclusive
    nrows = int(csum[-1].item()) if nb else 0
    # owners: rank o owns bins [edges[o], edges[o+1]) with about nrows / G rows each
    targets = torch.tensor([(o * nrows) // G for o in range(1, G)], dtype=torch.int64, device=csum.device)
    inner = torch.searchsorted(csum, targets, right=False) if G > 1 else targets[:0]
    edges = [0] + [min(int(x), nb) for x in inner.cpu().tolist()] + [nb]
    for o in range(1, G + 1):
        edges[o] = max(edges[o], edges[o - 1])
    ifg64 = ifg.to(torch.int64)
    lstart = [int(ifg64[edges[o]].item()) if edges[o] < nb else int(ig.numel()) for o in range(G)] + [int(ig.numel())]
    send_counts = [lstart[o + 1] - lstart[o] for o in range(G)]
    sc = ops.tensor(send_counts, dtype=torch.int64)
    rcv = torch.empty_like(sc)
    tdist.all_to_all_single(rcv, sc, group=group)
    recv_counts = rcv.cpu().tolist()
    got = _all_to_all_var(ig.to(torch.int64) + grouping.row_offset, send_counts, recv_counts, group)
    b0, b1 = edges[rank], edges[rank + 1]
    owned = torch.empty(int(sum(recv_counts)), dtype=torch.int64, device=got.device)
    base = gifg[b0:b1] - (gifg[b0] if b0 < nb else 0)
    pos = 0
    before = torch.zeros(b1 - b0, dtype=torch.int64, device=got.device)
    for r in range(G):
        c = cnts[r, b0:b1]
        m = int(recv_counts[r])
        if m:
            start = base + before                       # where rank r's rows of each owned bin begin
            excl = torch.cumsum(c, 0) - c
            dest = torch.repeat_interleave(start - excl, c) + torch.arange(m, device=got.device)
            owned[dest] = got[pos: pos + m]
        before = before + c
        pos += m
    return {"nCountGroup": total, "iFirstGroup": gifg, "bin_lo": b0, "bin_hi": b1, "iGroup_owned": owned}


SAMPLES_PER_RANK = 256


def _all_to_all_var(t: torch.Tensor, send_counts, recv_counts, group):
    """all-to-all of a 1-D tensor with per-destination element counts."""
    out = t.new_empty(int(sum(recv_counts)))
    tdist.all_to_all_single(out, t.contiguous(), output_split_sizes=[int(x) for x in recv_counts],
                            input_split_sizes=[int(x) for x in send_counts], group=group)
    return out


def sharded_lexsort(ops, key_cols, ascending=True, group=None):
    """rt.lexsort over row-sharded keys by sample sort: returns this rank's slice of the globally sorted
    permutation as int64 GLOBAL row ids — concatenated in rank order the slices equal np.lexsort of the concatenated
    keys (stable: ties are broken by global row id).  Steps: sample the shards, all-gather the samples, every rank
    picks the same G-1 splitters, rows are routed by `rtb_lex_partition`, bucketed (rc.BinCount pack), exchanged with
    one all-to-all per column, and the received rows are sorted locally with the global row id as the last key."""
    return _sample_sort(ops, key_cols, ascending, group, False)[0]


def _sample_sort(ops, key_cols, ascending, group, want_cols):
    """-> (sorted global row ids of this rank's slice, the key columns in that order (want_cols), row offsets, shard sizes)"""
    G, rank = _world(group)
    cols = [ops.as_col(c) for c in key_cols]
    n = ops.nrows(cols[0])
    asc = [bool(ascending)] * len(cols) if isinstance(ascending, (bool, np.bool_)) else [bool(a) for a in ascending]
    meta = ops.tensor([n], dtype=torch.int64)
    alln = meta.new_empty(G)
    tdist.all_gather_into_tensor(alln, meta, group=group)
    ns = alln.cpu().tolist()
    row_off = [0] * G
    for r in range(1, G):
        row_off[r] = row_off[r - 1] + ns[r - 1]
    dev = meta.device
    if G == 1:
        perm = ops.to_torch(ops.lexsort(cols, asc)).to(torch.int64)
        return perm, ([ops.take_rows(c, perm) for c in cols] if want_cols else None), row_off, ns

    # 1. samples -> splitters (identical on every rank)
    s_n = min(n, SAMPLES_PER_RANK)
    sidx = (torch.arange(s_n, device=dev, dtype=torch.int64) * (n // s_n)) if s_n else torch.zeros(0, dtype=torch.int64, device=dev)
    s_counts = [min(x, SAMPLES_PER_RANK) for x in ns]
    total_s = sum(s_counts)
    samples = []
    for c in cols:
        dt = ops.col_dtype(c)
        part = ops.col_bytes(ops.take_rows(c, sidx)) if s_n else ops.tensor([0], dtype=torch.uint8).new_empty(0)
        allb = _all_gather_var(part, [x * dt.itemsize for x in s_counts], group)
        samples.append(ops.col_from_bytes(allb, dt, total_s))
    s_rows = _all_gather_var(sidx + row_off[rank], s_counts, group)
    sperm = ops.to_torch(ops.lexsort([ops.rows_col(s_rows)] + samples, [True] + asc)).to(torch.int64)
    pick = sperm[[(k * total_s) // G for k in range(1, G)]] if total_s else sperm[:0]
    nsplit = int(pick.numel())
    splitters = [ops.take_rows(c, pick) for c in samples]
    split_rows = s_rows[pick]

    # 2. route and bucket the local rows
    if n:
        dest = ops.to_torch(ops.lex_partition(cols, splitters, split_rows, row_off[rank], asc)) if nsplit else \
            torch.zeros(n, dtype=torch.int32, device=dev)
        cnt, ig, _ = (ops.to_torch(x) for x in ops.pack(dest, G))
        send_counts = cnt.to(torch.int64).cpu().tolist()
        ig = ig.to(torch.int64)
    else:
        send_counts, ig = [0] * G, torch.zeros(0, dtype=torch.int64, device=dev)
    sc = ops.tensor(send_counts, dtype=torch.int64)
    rc_ = torch.empty_like(sc)
    tdist.all_to_all_single(rc_, sc, group=group)
    recv_counts = rc_.cpu().tolist()
    m = int(sum(recv_counts))

    # 3. exchange the key columns and the global row ids
    recv_cols = []
    for c in cols:
        dt = ops.col_dtype(c)
        payload = ops.col_bytes(ops.take_rows(c, ig)) if n else ops.tensor([0], dtype=torch.uint8).new_empty(0)
        got = _all_to_all_var(payload, [x * dt.itemsize for x in send_counts], [x * dt.itemsize for x in recv_counts], group)
        recv_cols.append(ops.col_from_bytes(got, dt, m))
    recv_rows = _all_to_all_var(ig + row_off[rank], send_counts, recv_counts, group)

    # 4. local sort of what arrived; the global row id is the least significant key
    if m == 0:
        return recv_rows, (recv_cols if want_cols else None), row_off, ns
    perm = ops.to_torch(ops.lexsort([ops.rows_col(recv_rows)] + recv_cols, [True] + asc)).to(torch.int64)
    return recv_rows[perm], ([ops.take_rows(c, perm) for c in recv_cols] if want_cols else None), row_off, ns


def sharded_group_from_lexsort(ops, key_cols, ascending=True, group=None, replicate=True):
    """rc.GroupFromLexSort over row-sharded keys: sample-sort the rows, flag the bin boundaries inside
    each rank's sorted slice, settle the one boundary between neighbouring slices by exchanging their first / last key
    rows, number the bins by an exclusive scan of the per-rank distinct counts, and send every (row, bin) pair back to
    the rank that owns the row.  Bins are numbered in sorted order, base 1, exactly as the single-GPU call numbers them.
    -> dict: ikey (int32, this rank's rows), unique_count, igroup_slice (this rank's slice of the sorted global row
    ids), bin_lo / nCountGroup_owned / iFirstKey_owned (the bins that START in this rank's slice: global bin ids
    bin_lo+1 ...), and with replicate=True the full nCountGroup [U+1] and iFirstKey [U] (global row ids)."""
    G, rank = _world(group)
    rows, scols, row_off, ns = _sample_sort(ops, key_cols, ascending, group, True)
    m = int(rows.numel())
    dev = rows.device
    n_local = ns[rank]

    # 1. bins inside the slice (identity iGroup: the columns are already in sorted order)
    if m:
        lbin, lfirst, lcount = (ops.to_torch(x) for x in ops.group_from_sorted(scols))
        lbin, lfirst, lcount = lbin.to(torch.int64), lfirst.to(torch.int64), lcount.to(torch.int64)
        uloc = int(lfirst.numel())
        ends = torch.tensor([0, m - 1], dtype=torch.int64, device=dev)
        edge = torch.cat([ops.col_bytes(ops.take_rows(c, ends)).reshape(2, -1) for c in scols], dim=1).contiguous()
        first_cnt, last_cnt = int(lcount[1]), int(lcount[uloc])
    else:
        uloc, first_cnt, last_cnt = 0, 0, 0
        edge = None
    row_bytes = sum(ops.col_dtype(c).itemsize for c in scols)

    # 2. neighbour boundaries: every rank sees every slice's (size, local bins, edge counts, first / last key row)
    meta = ops.tensor([m, uloc, first_cnt, last_cnt], dtype=torch.int64)
    allmeta = meta.new_empty(4 * G)
    tdist.all_gather_into_tensor(allmeta, meta, group=group)
    allmeta = allmeta.cpu().view(G, 4).tolist()
    payload = edge.reshape(-1) if m else ops.tensor([0], dtype=torch.uint8).new_empty(0)
    alledge = _all_gather_var(payload, [2 * row_bytes if x[0] else 0 for x in allmeta], group).cpu()
    edges, pos = [], 0
    for x in allmeta:
        if x[0]:
            edges.append((alledge[pos: pos + row_bytes], alledge[pos + row_bytes: pos + 2 * row_bytes]))
            pos += 2 * row_bytes
        else:
            edges.append(None)
    cont, prev = [0] * G, None
    for r in range(G):
        if edges[r] is None:
            continue
        if prev is not None and torch.equal(edges[prev][1], edges[r][0]):
            cont[r] = 1  # this slice opens inside the bin the previous slice closed in
        prev = r
    distinct = [allmeta[r][1] - cont[r] for r in range(G)]
    offset = [0] * G
    for r in range(1, G):
        offset[r] = offset[r - 1] + distinct[r - 1]
    U = offset[-1] + distinct[-1]

    # 3. (row, bin) back to the owner of the row
    ikey = torch.zeros(max(n_local, 1), dtype=torch.int32, device=dev)[:n_local]
    if m:
        gbin = (lbin + (offset[rank] - cont[rank])).to(torch.int32)
        bounds = torch.tensor([row_off[r] + ns[r] for r in range(G - 1)], dtype=torch.int64, device=dev)
        dest = torch.bucketize(rows, bounds, right=True).to(torch.int32) if G > 1 else torch.zeros(m, dtype=torch.int32, device=dev)
        cnt, ig, _ = (ops.to_torch(x) for x in ops.pack(dest, G))
        send_counts = cnt.to(torch.int64).cpu().tolist()
        ig = ig.to(torch.int64)
        send_rows, send_bins = rows[ig], gbin[ig]
    else:
        send_counts = [0] * G
        send_rows, send_bins = rows, torch.zeros(0, dtype=torch.int32, device=dev)
    if G > 1:
        sc = ops.tensor(send_counts, dtype=torch.int64)
        rc_ = torch.empty_like(sc)
        tdist.all_to_all_single(rc_, sc, group=group)
        recv_counts = rc_.cpu().tolist()
        got_rows = _all_to_all_var(send_rows, send_counts, recv_counts, group)
        got_bins = _all_to_all_var(send_bins, send_counts, recv_counts, group)
    else:
        got_rows, got_bins = send_rows, send_bins
    if got_rows.numel():
        ikey[got_rows - row_off[rank]] = got_bins

    # 4. the bins that start here: counts (a bin may run on through the following slices) and first rows
    if m and distinct[rank] > 0:
        own_cnt = lcount[1 + cont[rank]:].clone()
        q = rank + 1
        extra = 0
        while q < G:
            if allmeta[q][0] == 0:
                q += 1
                continue
            if not cont[q]:
                break
            extra += allmeta[q][2]
            if allmeta[q][1] > 1:
                break
            q += 1
        own_cnt[-1] += extra
        own_first = rows[lfirst[cont[rank]:]]
    else:
        own_cnt = torch.zeros(0, dtype=torch.int64, device=dev)
        own_first = torch.zeros(0, dtype=torch.int64, device=dev)
    out = {"ikey": ikey, "unique_count": int(U), "igroup_slice": rows, "bin_lo": offset[rank],
           "nCountGroup_owned": own_cnt.to(torch.int32), "iFirstKey_owned": own_first}
    if replicate:
        allc = _all_gather_var(own_cnt.to(torch.int32), distinct, group) if G > 1 else own_cnt.to(torch.int32)
        allf = _all_gather_var(own_first, distinct, group) if G > 1 else own_first
        out["nCountGroup"] = torch.cat([allc.new_zeros(1), allc])
        out["iFirstKey"] = allf
    return out


class HostFeed:
    """A shard that lives in (pinned) host memory, on its way to the device while it is being grouped.

    All host -> device copies are enqueued up front on their own stream, chunk by chunk, key and value of a chunk
    together; the grouping waits for a chunk's event right before it pushes the chunk, and every chunk's iKey goes back
    to the host on a third stream as soon as it is final -- PCIe runs in both directions under the hash passes.
    The first chunk is the seeded path's prefix (SEED_ROWS)."""

    def __init__(self, key_host: np.ndarray, val_host: np.ndarray, device=None, chunk_rows=None):
        n = len(key_host)
        if len(val_host) != n:
            raise ValueError("key and value columns differ in length")
        self.n = n
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.chunk_rows = int(chunk_rows or HOST_CHUNK_ROWS)
        first = (min(n, SEED_ROWS) & ~127) if n > SEED_ROWS else n
        self.bounds = [0]
        if first:
            self.bounds.append(first)
        self.bounds += [hi for _, hi in self.chunks(first, n)]
        kt, vt = torch.from_numpy(key_host), torch.from_numpy(val_host)
        self.key = torch.empty(n, dtype=kt.dtype, device=self.device)
        self.val = torch.empty(n, dtype=vt.dtype, device=self.device)
        self.host_ikey = torch.empty(n, dtype=torch.int32, pin_memory=True)
        cur = torch.cuda.current_stream(self.device)
        self.up, self.down = torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)
        self.up.wait_stream(cur)
        self.events = []
        from . import rc as _rc  # pinned arrays are copied asynchronously, pageable ones through the shim's staging ring

        for lo, hi in zip(self.bounds[:-1], self.bounds[1:]):
            _rc._h2d_chunk(self.key[lo:hi].view(torch.uint8), key_host[lo:hi], self.up)
            _rc._h2d_chunk(self.val[lo:hi].view(torch.uint8), val_host[lo:hi], self.up)
            ev = torch.cuda.Event()
            ev.record(self.up)
            self.events.append((hi, ev))
        self.h2d_bytes = n * (kt.element_size() + vt.element_size())

    def chunks(self, lo, hi):
        c = self.chunk_rows
        return [(a, min(hi, a + c)) for a in range(lo, hi, c)]

    def wait_rows(self, hi):
        """The current stream waits until rows [0, hi) are on the device."""
        cur = torch.cuda.current_stream(self.device)
        for end, ev in self.events:
            cur.wait_event(ev)
            if end >= hi:
                break

    def ikey_final(self, lo, hi, ikey):
        """iKey[lo:hi] holds its final values on the current stream: start its copy to the host."""
        self.down.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.down):
            self.host_ikey[lo:hi].copy_(ikey[lo:hi], non_blocking=True)

    def finish(self, ikey=None):
        """-> the host iKey (NumPy).  `ikey`: re-read everything from this device tensor (the streamed copies are stale)."""
        if ikey is not None:
            self.down.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.down):
                self.host_ikey.copy_(ikey, non_blocking=True)
        self.down.synchronize()
        torch.cuda.current_stream(self.device).wait_stream(self.up)
        return self.host_ikey.numpy()


def sharded_groupby_sum_host(key_host, values_host, group=None, ops=None, shard_sizes=None, chunk_rows=None):
    """sharded_groupby_sum for shards held in host memory (NumPy in, NumPy out): every rank streams its shard to its GPU
    while the shard is being grouped and streams iKey back (HostFeed).  Pinned inputs make the copies asynchronous.
    -> (iKey int32[n] of this shard, iFirstKey int64[U] global rows, U, sums[U + 1])."""
    ops = ops or CudaOps()
    feed = HostFeed(np.ascontiguousarray(key_host), np.ascontiguousarray(values_host), ops.device, chunk_rows)
    g = sharded_groupby(ops, [feed.key], None, group, sum_values=feed.val, shard_sizes=shard_sizes, feed=feed)
    if g.partial_sums is not None:
        s = _combine("sum", {"v": g.partial_sums[: g.unique_count + 1]}, g.unique_count + 1, False, group, ALLTOALL_MIN_BINS)["v"]
        if feed.val.dtype == torch.float32:
            s = s.to(torch.float32)
    else:
        s = sharded_reduce(ops, feed.val, g, GB.GB_SUM, group)
    ik = feed.finish(None if g.host_ikey_streamed else g.ikey[: feed.n])
    return ik, g.ifirstkey.cpu().numpy(), g.unique_count, s.cpu().numpy()


def sharded_groupby_sum(key_cols, values, group=None, ops=None, shard_sizes=None):
    """bench.py's multi-GPU step: sharded hash grouping + GB_SUM.  -> (ikey, ifirstkey, U, sums[U+1])."""
    ops = ops or CudaOps()
    g = sharded_groupby(ops, key_cols, None, group, sum_values=values, shard_sizes=shard_sizes)
    if g.partial_sums is not None:
        # the grouping pass summed the shard as it went (fused kernel) and its bins are global: one all-reduce is left
        s = _combine("sum", {"v": g.partial_sums[: g.unique_count + 1]}, g.unique_count + 1, False, group, ALLTOALL_MIN_BINS)["v"]
        vcol = ops.as_col(values)
        if ops.is_f32(vcol):
            s = s.to(torch.float32)
        return g.ikey, g.ifirstkey, g.unique_count, s
    s = sharded_reduce(ops, values, g, GB.GB_SUM, group)  # runs while the re-index pass is still on its side stream
    return g.ikey, g.ifirstkey, g.unique_count, s
