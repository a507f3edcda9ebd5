"""World-size-2 gloo run of riptable_b200.dist on CPU: the shard-merge logic and every collective it issues
(all-gather of uniques, all-reduce / all-gather / all-to-all of partials) against a single-process oracle run.
The device ops are replaced by an oracle-backed stand-in (tests only); the product backend is CudaOps."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import riptide_oracle as orc


def _np(x):
    return x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class _StreamOverflow(RuntimeError):
    """Stand-in for RtbWorkspaceError: the resumable grouping ran out of provisioned bins."""


class OracleOps:
    """NumPy/oracle stand-in for CudaOps (same method set), CPU tensors for the collectives."""

    workspace_error = (_StreamOverflow,)

    def as_col(self, x):
        return x.numpy() if isinstance(x, torch.Tensor) else np.ascontiguousarray(x)

    def nrows(self, col):
        return len(col)

    def take_rows(self, col, idx):
        return np.ascontiguousarray(_np(col)[_np(idx)])

    def col_bytes(self, col):
        return torch.from_numpy(np.ascontiguousarray(col).view(np.uint8).reshape(-1).copy())

    def col_from_bytes(self, t, dtype, n):
        return t.numpy().view(np.dtype(dtype))[:n].copy() if n else np.zeros(0, np.dtype(dtype))

    def col_dtype(self, col):
        return col.dtype

    def groupby(self, cols, filter=None):
        return orc.MultiKeyGroupBy32(list(cols), 0, filter, 2)

    def remap(self, ikey, table):
        return table[ikey.to(torch.int64)], None

    def wait(self, ev):
        pass

    def torch_dtype(self, col):
        return torch.from_numpy(np.asarray(col)[:0].copy()).dtype

    def reduce(self, val, ikey, nuniq, func):
        ik = ikey.numpy() if isinstance(ikey, torch.Tensor) else ikey
        return orc.GroupByAll32([np.asarray(val)], ik, nuniq, [int(func)], [0], [nuniq + 1], 0)[0]

    def bincount(self, ikey, nbins):
        ik = ikey.numpy() if isinstance(ikey, torch.Tensor) else ikey
        return orc.BinCount(ik, nbins)

    def valid_mask_as_i8(self, col):
        return orc._is_valid_mask(np.asarray(col)).astype(np.int8)

    def ifirst(self, ikey, nuniq, mode):
        return orc.MakeiFirst(_np(ikey), nuniq, None, mode)

    def pack(self, ikey, nbins):
        return orc.BinCount(_np(ikey), nbins, pack=True)

    def cumsum(self, val, ikey, nuniq, filter):
        return orc.EmaAll32([_np(val)], _np(ikey), nuniq, 300, (0.0, None, None if filter is None else _np(filter), None))[0]

    def combine_filter(self, ikey, filter):
        return orc.CombineFilter(_np(ikey), _np(filter))

    def as_f64(self, col):
        return _np(col).astype(np.float64)

    def lexsort(self, cols, ascending=True):
        return orc.LexSort32([_np(c) for c in cols], ascending=ascending)

    def lex_partition(self, cols, splitters, splitter_rows, row_offset, ascending):
        # rank every row among the splitters with one lexsort of [rows ; splitters] (global row id = last tie-break)
        cols = [_np(c) for c in cols]
        spl = [_np(c) for c in splitters]
        n, k = len(cols[0]), len(spl[0])
        grow = np.concatenate([np.arange(n, dtype=np.int64) + row_offset, _np(splitter_rows).astype(np.int64)])
        keys = [grow] + [np.concatenate([c, sc]) for c, sc in zip(cols, spl)]
        perm = orc.LexSort32(keys, ascending=[True] + list(ascending))
        is_split = perm >= n
        before = np.cumsum(is_split) - is_split  # splitters strictly before each sorted position
        dest = np.empty(n + k, np.int32)
        dest[perm] = before
        # a row equal to a splitter IS that splitter's row: it counts as "at or before"
        out = dest[:n].copy()
        srow = _np(splitter_rows).astype(np.int64)
        for j in range(k):
            loc = srow[j] - row_offset
            if 0 <= loc < n:
                out[loc] += 1
        return out

    def rows_col(self, rows_i64):
        return _np(rows_i64).astype(np.int64)

    def group_from_sorted(self, cols):
        cols = [_np(c) for c in cols]
        ik, ifk, cnt = orc.GroupFromLexSort(np.arange(len(cols[0]), dtype=np.int32), cols)
        return torch.from_numpy(ik.copy()), torch.from_numpy(ifk.copy()), torch.from_numpy(cnt.copy())

    def is_f32(self, col):
        return _np(col).dtype == np.float32

    def invalid_item_bytes(self, col):
        dt = _np(col).dtype
        if dt.kind in "SUV":
            return torch.zeros(dt.itemsize, dtype=torch.uint8)
        return torch.from_numpy(np.array([orc.invalid_for(dt)], dtype=dt).view(np.uint8).copy())

    def tensor(self, x, dtype=None):
        return torch.as_tensor(x, dtype=dtype)

    # --- resumable single-key grouping (stand-in for rc.GroupByStream) ---------------------------------------
    def stream(self, max_unique, max_chunk_rows, with_sums=False):
        return _OracleStream(max_unique, with_sums)

    def can_stream(self, col):
        return col.dtype.kind in "biuf" and col.dtype.itemsize in (1, 2, 4, 8)

    def bin_table(self, newbins):
        return torch.cat([torch.zeros(1, dtype=torch.int32), newbins.to(torch.int32)])

    def stream_fix_first(self, st, u0, u_mid, old_first):
        st.first[u0:u_mid] = _np(old_first)[st.first[u0:u_mid].astype(np.int64)].astype(np.int32)

    def remap_prefix(self, ikey, S, table):
        if S:
            ikey[:S] = table[ikey[:S].to(torch.int64)]
        return ikey

    def stream_move_sums(self, old, new, table, u_loc):
        acc = np.zeros(max(len(new.acc), int(table.max()) + 1), dtype=old.acc.dtype)
        acc[: len(new.acc)] = new.acc.astype(old.acc.dtype)
        acc[_np(table).astype(np.int64)] = old.acc[: u_loc + 1]
        new.acc = acc

    def slice_rows(self, col, lo, hi):
        return col[lo:hi]

    def slice_filter(self, f, lo, hi):
        return None if f is None else np.asarray(f)[lo:hi]

    def empty_ikey(self, n):
        return torch.zeros(max(n, 4), dtype=torch.int32)

    def to_torch(self, x):
        return torch.from_numpy(np.ascontiguousarray(x)) if isinstance(x, np.ndarray) else x


class _OracleStream:
    """Chunked first-appearance grouping restated with the oracle: group [uniques so far ; chunk] in one call."""

    def __init__(self, max_unique=1 << 62, with_sums=False):
        self.uniq = None
        self.max_unique = max_unique
        self.with_sums = with_sums
        self.acc = np.zeros(1)
        self.first = np.zeros(0, np.int32)

    def sums(self):
        return torch.from_numpy(self.acc.copy())

    @property
    def unique_count(self):
        return len(self.first)

    @property
    def ifirstkey(self):
        return self.first

    def push(self, keys, filter=None, row_offset=0, out=None, values=None):
        keys = np.asarray(keys)
        u = self.unique_count
        allk = keys if self.uniq is None else np.concatenate([self.uniq, keys])
        f = None
        if filter is not None:
            f = np.concatenate([np.ones(u, bool), np.asarray(filter, dtype=bool)])
        ik, ifk, U = orc.MultiKeyGroupBy32([allk], 0, f, 2)
        assert np.array_equal(ik[:u], np.arange(1, u + 1))
        if U > self.max_unique:
            raise _StreamOverflow(f"more than {self.max_unique} unique keys")
        self.first = np.concatenate([self.first, (ifk[u:] - u + row_offset).astype(np.int32)])
        self.uniq = allk[ifk]
        res = torch.from_numpy(ik[u:].copy())
        if self.with_sums:
            acc = np.zeros(U + 1, dtype=self.acc.dtype if values is None else (np.float64 if np.asarray(values).dtype.kind == "f" else np.int64))
            acc[: len(self.acc)] = self.acc
            if values is not None:
                part = orc.GroupByAll32([np.asarray(values)], ik[u:], U, [0], [1], [U + 1], 0)[0]
                acc += part.astype(acc.dtype)
            self.acc = acc
        if out is not None:
            out[: len(res)] = res
        return res



def _make(seed, n):
    rng = np.random.default_rng(seed)
    k1 = rng.integers(0, 40, n).astype(np.int64) * 1_000_003
    k2 = rng.choice(np.array([b"aa", b"b", b"", b"cccc"], dtype="S4"), n)
    v = rng.random(n) * 100 - 50
    v[rng.random(n) < 0.1] = np.nan
    vi = rng.integers(-1000, 1000, n).astype(np.int32)
    vi[rng.random(n) < 0.05] = np.iinfo(np.int32).min
    f = rng.random(n) < 0.8
    return k1, k2, v, vi, f


def _worker(rank, world, port, n, cuts, a2a, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from riptable_b200 import dist as rd

        rd.SEED_ROWS = 1024  # shard 0's seed prefix is much smaller than the shard: late keys get merged
        k1, k2, v, vi, f = _make(7, n)
        lo, hi = cuts[rank], cuts[rank + 1]
        ops = OracleOps()
        out = {}
        for name, keys, filt in (("one", [k1], None), ("onef", [k1], f), ("two", [k1, k2], None), ("filt", [k1, k2], f)):
            g = rd.sharded_groupby(ops, [k[lo:hi] for k in keys], None if filt is None else filt[lo:hi])
            res = {"ikey": g.ikey.numpy().copy(), "ifk": g.ifirstkey.numpy().copy(), "u": g.unique_count,
                   "off": g.row_offset}
            for fn in (0, 1, 2, 3, 4, 5, 50, 51, 52, 53, 54, 55):
                res[f"f{fn}"] = rd.sharded_reduce(ops, v[lo:hi], g, fn, alltoall_min_bins=a2a).numpy().copy()
            for fn in (0, 2, 3, 52, 53):
                res[f"i{fn}"] = rd.sharded_reduce(ops, vi[lo:hi], g, fn, alltoall_min_bins=a2a).numpy().copy()
            res["count"] = rd.sharded_count(ops, g, alltoall_min_bins=a2a).numpy().copy()
            for fn, p in ((100, 0), (102, 0), (101, 0), (101, 3), (101, -1), (101, -4), (101, 10_000)):
                res[f"pick{fn}_{p}"] = np.asarray(rd.sharded_pick(ops, v[lo:hi], g, fn, p)).copy()
                res[f"picks{fn}_{p}"] = np.asarray(rd.sharded_pick(ops, k2[lo:hi], g, fn, p)).copy()
            if name in ("one", "filt"):
                pk = rd.sharded_pack(ops, g)
                res["pack"] = {k2_: (_np(v2_).copy() if hasattr(v2_, "shape") else v2_) for k2_, v2_ in pk.items()}
            if filt is None:
                gi, gf, gu, gs = rd.sharded_groupby_sum([k[lo:hi] for k in keys], np.nan_to_num(v[lo:hi]), ops=ops)
                res["gbsum"] = (_np(gi).copy(), _np(gf).copy(), gu, _np(gs).copy())
                gi, gf, gu, gs = rd.sharded_groupby_sum([k[lo:hi] for k in keys], vi[lo:hi], ops=ops)
                res["gbsum_i"] = (_np(gi).copy(), _np(gs).copy())
            res["cumsum_v"] = _np(rd.sharded_cumsum(ops, np.nan_to_num(v[lo:hi]), g)).copy()
            res["cumsum_vi"] = _np(rd.sharded_cumsum(ops, vi[lo:hi], g, f[lo:hi])).copy()
            res["cumsum_f32"] = _np(rd.sharded_cumsum(ops, np.nan_to_num(v[lo:hi]).astype(np.float32), g, f[lo:hi])).copy()
            out[name] = res
        q.put((rank, out))
    finally:
        tdist.destroy_process_group()


def _sort_worker(rank, world, port, cuts, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from riptable_b200 import dist as rd

        rd.SAMPLES_PER_RANK = 32
        a, f, s = _sort_data(cuts[-1])
        lo, hi = cuts[rank], cuts[rank + 1]
        ops = OracleOps()
        out = {}
        out["afs"] = _np(rd.sharded_lexsort(ops, [a[lo:hi], f[lo:hi], s[lo:hi]])).copy()
        out["desc"] = _np(rd.sharded_lexsort(ops, [f[lo:hi], a[lo:hi]], ascending=[False, True])).copy()
        out["one"] = _np(rd.sharded_lexsort(ops, [a[lo:hi]])).copy()
        for name, cols in (("g_afs", [a[lo:hi], f[lo:hi], s[lo:hi]]), ("g_one", [a[lo:hi]]), ("g_const", [np.zeros(hi - lo, np.int8)])):
            g = rd.sharded_group_from_lexsort(ops, cols)
            out[name] = {k: (_np(v).copy() if isinstance(v, torch.Tensor) else v) for k, v in g.items()}
        q.put((rank, out))
    finally:
        tdist.destroy_process_group()


def _sort_data(n):
    rng = np.random.default_rng(11)
    a = rng.integers(-5, 5, n).astype(np.int16)
    f = np.round(rng.random(n) * 6, 0)
    f[rng.random(n) < 0.1] = np.nan
    f[rng.random(n) < 0.1] = -0.0
    s = rng.choice(np.array([b"", b"a", b"ab", b"zz"], dtype="S3"), n)
    return a, f, s


@pytest.mark.parametrize("cuts", [(0, 2500, 6001), (0, 0, 3001), (0, 5000, 5003)])
def test_sharded_lexsort_world2(cuts):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sort_worker, args=(r, 2, port, cuts, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, f, s = _sort_data(cuts[-1])
    np.testing.assert_array_equal(np.concatenate([got[0]["afs"], got[1]["afs"]]), np.lexsort([a, f, s]))
    np.testing.assert_array_equal(np.concatenate([got[0]["desc"], got[1]["desc"]]),
                                  orc.LexSort32([f, a], ascending=[False, True]))
    np.testing.assert_array_equal(np.concatenate([got[0]["one"], got[1]["one"]]), np.lexsort([a]))
    assert len(got[0]["afs"]) > 0 and len(got[1]["afs"]) > 0  # the splitter really splits the rows
    # GroupFromLexSort across the ranks == the single-process call on the concatenated columns
    for name, cols in (("g_afs", [a, f, s]), ("g_one", [a]), ("g_const", [np.zeros(len(a), np.int8)])):
        eik, eifk, ecnt = orc.GroupFromLexSort(orc.LexSort32(cols), cols)
        g0, g1 = got[0][name], got[1][name]
        np.testing.assert_array_equal(np.concatenate([g0["ikey"], g1["ikey"]]), eik)
        assert g0["unique_count"] == g1["unique_count"] == len(eifk)
        for g in (g0, g1):
            np.testing.assert_array_equal(g["iFirstKey"], eifk)
            np.testing.assert_array_equal(g["nCountGroup"], ecnt)
        np.testing.assert_array_equal(np.concatenate([g0["iFirstKey_owned"], g1["iFirstKey_owned"]]), eifk)
        np.testing.assert_array_equal(np.concatenate([g0["nCountGroup_owned"], g1["nCountGroup_owned"]]), ecnt[1:])
        assert g0["bin_lo"] == 0 and g1["bin_lo"] == len(g0["iFirstKey_owned"])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("a2a", [1 << 40, 1])  # all-reduce / all-gather merge, then the all-to-all merge
@pytest.mark.parametrize("cuts", [(0, 3000, 7001), (0, 0, 7001), (0, 7000, 7001)])
def test_sharded_groupby_and_reduce_world2(cuts, a2a):
    n = cuts[-1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, cuts, a2a, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    k1, k2, v, vi, f = _make(7, n)
    for name, keys, filt in (("one", [k1], None), ("onef", [k1], f), ("two", [k1, k2], None), ("filt", [k1, k2], f)):
        eik, eifk, eu = orc.MultiKeyGroupBy32(keys, 0, filt, 2)
        full = np.concatenate([got[r][name]["ikey"] for r in range(2)])
        np.testing.assert_array_equal(full, eik)               # global bin ids, bit-exact with the single-process run
        for r in range(2):
            assert got[r][name]["u"] == eu
            np.testing.assert_array_equal(got[r][name]["ifk"], eifk)
            assert got[r][name]["off"] == cuts[r]
            np.testing.assert_array_equal(got[r][name]["count"], orc.BinCount(eik, eu + 1))
            for fn in (0, 1, 2, 3, 4, 5, 50, 51, 52, 53, 54, 55):
                exp = orc.GroupByAll32([v], eik, eu, [fn], [0], [eu + 2], 0)[0]
                np.testing.assert_allclose(got[r][name][f"f{fn}"], exp, rtol=1e-11, atol=1e-9, equal_nan=True,
                                           err_msg=f"{name} func {fn}")
            for fn in (0, 2, 3, 52, 53):
                exp = orc.GroupByAll32([vi], eik, eu, [fn], [0], [eu + 2], 0)[0]
                np.testing.assert_array_equal(got[r][name][f"i{fn}"], exp, err_msg=f"{name} int func {fn}")
            cnt, ig, ifg = orc.BinCount(eik, eu + 1, pack=True)
            for fn, p in ((100, 0), (102, 0), (101, 0), (101, 3), (101, -1), (101, -4), (101, 10_000)):
                exp = orc.GroupByAllPack32([v], eik, ig, ifg, cnt, eu, [fn], [1], [eu + 1], p)[0]
                np.testing.assert_array_equal(got[r][name][f"pick{fn}_{p}"][1:].view(np.uint64), exp[1:].view(np.uint64),
                                              err_msg=f"{name} pick {fn} {p}")
                exp = orc.GroupByAllPack32([k2], eik, ig, ifg, cnt, eu, [fn], [1], [eu + 1], p)[0]
                np.testing.assert_array_equal(got[r][name][f"picks{fn}_{p}"][1:], exp[1:], err_msg=f"{name} pick S {fn} {p}")
        if name in ("one", "filt"):
            for r in range(2):
                pk = got[r][name]["pack"]
                np.testing.assert_array_equal(pk["nCountGroup"], cnt)
                np.testing.assert_array_equal(pk["iFirstGroup"], ifg)
                b0, b1 = pk["bin_lo"], pk["bin_hi"]
                lo_p = int(ifg[b0]) if b0 < len(ifg) else len(ig)
                hi_p = int(ifg[b1]) if b1 < len(ifg) else len(ig)
                np.testing.assert_array_equal(pk["iGroup_owned"], ig[lo_p:hi_p], err_msg=f"{name} pack rank {r}")
            assert got[0][name]["pack"]["bin_hi"] == got[1][name]["pack"]["bin_lo"]
        if filt is None:
            np.testing.assert_array_equal(np.concatenate([got[r][name]["gbsum"][0] for r in range(2)]), eik)
            np.testing.assert_array_equal(np.concatenate([got[r][name]["gbsum_i"][0] for r in range(2)]), eik)
            for r in range(2):
                np.testing.assert_array_equal(got[r][name]["gbsum"][1], eifk)
                assert got[r][name]["gbsum"][2] == eu
                exp = orc.GroupByAll32([np.nan_to_num(v)], eik, eu, [0], [1], [eu + 1], 0)[0]
                np.testing.assert_allclose(got[r][name]["gbsum"][3][1:], exp[1:], rtol=1e-12, atol=1e-9)
                exp = orc.GroupByAll32([vi], eik, eu, [0], [1], [eu + 1], 0)[0]
                np.testing.assert_array_equal(got[r][name]["gbsum_i"][1][1:], exp[1:])
        lo_, hi_ = cuts[0], cuts[2]
        full = lambda key: np.concatenate([got[r][name][key] for r in range(2)])  # noqa: E731
        exp = orc.EmaAll32([np.nan_to_num(v)], eik, eu, 300, (0.0, None, None, None))[0]
        np.testing.assert_allclose(full("cumsum_v"), exp, rtol=1e-12, atol=1e-9, equal_nan=True)
        exp = orc.EmaAll32([vi], eik, eu, 300, (0.0, None, f, None))[0]
        np.testing.assert_array_equal(full("cumsum_vi"), exp)
        exp = orc.EmaAll32([np.nan_to_num(v).astype(np.float32)], eik, eu, 300, (0.0, None, f, None))[0]
        assert full("cumsum_f32").dtype == np.float32
        np.testing.assert_allclose(full("cumsum_f32"), exp, rtol=2e-5, atol=1e-2, equal_nan=True)


def _overflow_worker(rank, world, port, cuts, where, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from riptable_b200 import dist as rd

        rd.SEED_ROWS = 1024
        rd.SEEDED_MIN_UNIQUE = 64  # the seeded table holds 64 bins; one shard has far more keys than that
        k = _overflow_keys(cuts[-1], where, cuts)
        lo, hi = cuts[rank], cuts[rank + 1]
        g = rd.sharded_groupby(OracleOps(), [k[lo:hi]], None)
        q.put((rank, {"ikey": g.ikey.numpy().copy(), "ifk": g.ifirstkey.numpy().copy(), "u": g.unique_count}))
    finally:
        tdist.destroy_process_group()


def _overflow_keys(n, where, cuts):
    rng = np.random.default_rng(3)
    k = rng.integers(0, 30, n).astype(np.int64)
    lo, hi = cuts[where], cuts[where + 1]
    k[lo:hi] = rng.integers(0, 900, hi - lo)  # only this shard overflows
    return k


@pytest.mark.parametrize("where", [0, 1])
def test_seeded_overflow_on_one_rank_falls_back_on_all_ranks(where):
    """a shard with more uniques than the seeded table holds fails on that rank alone; every rank must
    leave the seeded path together (no rank may be left waiting in a collective) and the result stays exact."""
    cuts = (0, 3000, 6500)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_overflow_worker, args=(r, 2, port, cuts, where, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k = _overflow_keys(cuts[-1], where, cuts)
    eik, eifk, eu = orc.MultiKeyGroupBy32([k], 0, None, 2)
    assert eu > 64
    np.testing.assert_array_equal(np.concatenate([got[0]["ikey"], got[1]["ikey"]]), eik)
    for r in range(2):
        assert got[r]["u"] == eu
        np.testing.assert_array_equal(got[r]["ifk"], eifk)
