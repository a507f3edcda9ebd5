"""GPU, needs >= 2 devices (skipped on a single-GPU box): the real NCCL path of riptable_b200.dist on two ranks —
seeded grouping with late keys, reductions, first/last, cumsum and the sample-sort lexsort — against the oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import riptide_oracle as orc


def _data(n):
    rng = np.random.default_rng(5)
    k = rng.integers(0, 3000, n).astype(np.int64) * 7919
    k[: n // 2] = k[: n // 2] % (7919 * 1500)  # half of the keys only appear in the second shard
    f = np.round(rng.random(n) * 40, 0)
    f[rng.random(n) < 0.05] = np.nan
    v = rng.random(n) * 10 - 5
    s = rng.choice(np.array([b"", b"a", b"ab", b"zz"], dtype="S3"), n)
    return k, f, v, s


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as tdist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    tdist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from riptable_b200 import dist as rd

        rd.SEED_ROWS = 4096
        rd.SAMPLES_PER_RANK = 64
        k, f, v, s = _data(n)
        lo, hi = (0, n // 3) if rank == 0 else (n // 3, n)
        ops = rd.CudaOps()
        g = rd.sharded_groupby(ops, [torch.from_numpy(k[lo:hi]).cuda()], None)
        out = {"ikey": g.ikey.cpu().numpy(), "ifk": g.ifirstkey.cpu().numpy(), "u": g.unique_count}
        vt = torch.from_numpy(v[lo:hi]).cuda()
        for fn in (0, 1, 3, 5):
            out[f"r{fn}"] = rd.sharded_reduce(ops, vt, g, fn).cpu().numpy()
        out["first"] = rd.sharded_pick(ops, vt, g, 100).numpy()
        out["last"] = rd.sharded_pick(ops, vt, g, 102).numpy()
        out["cumsum"] = rd.sharded_cumsum(ops, vt, g).cpu().numpy()
        pk = rd.sharded_pack(ops, g)
        out["pack"] = {a: (b.cpu().numpy() if hasattr(b, "cpu") else b) for a, b in pk.items()}
        g2 = rd.sharded_groupby(ops, [torch.from_numpy(k[lo:hi]).cuda(), s[lo:hi]], None)
        out["ikey2"] = g2.ikey.cpu().numpy()
        out["sum2"] = rd.sharded_reduce(ops, vt, g2, 0).cpu().numpy()
        out["sort"] = rd.sharded_lexsort(ops, [torch.from_numpy(f[lo:hi]).cuda(), s[lo:hi], torch.from_numpy(k[lo:hi]).cuda()]).cpu().numpy()
        fs = [torch.from_numpy(f[lo:hi]).cuda(), s[lo:hi]]
        g = rd.sharded_group_from_lexsort(ops, fs)
        out["gfl"] = {k_: (v_.cpu().numpy() if isinstance(v_, torch.Tensor) else v_) for k_, v_ in g.items()}
        # the fused group-and-sum step of bench.py: with late keys (falls back to the merge + separate reduction), and
        # with every key inside shard 0's seed prefix (sums come out of the grouping pass, one all-reduce)
        kt = torch.from_numpy(k[lo:hi]).cuda()
        for name, keys, seed_rows in (("gs_late", kt, 4096), ("gs_seeded", kt % (7919 * 700), 1 << 25), ("gs_seeded_small_prefix", kt % (7919 * 700), 65_536)):
            rd.SEED_ROWS = seed_rows
            for vname, vals in (("f64", vt), ("f32", vt.float()), ("i32", (vt * 100).to(torch.int32))):
                gi, gf, gu, gs = rd.sharded_groupby_sum([keys], vals, ops=ops, shard_sizes=[n // 3, n - n // 3])
                out[f"{name}_{vname}"] = (gi.cpu().numpy(), gf.cpu().numpy(), gu, gs.cpu().numpy())
        # the same step for shards held in host memory (HostFeed: chunked copies under the grouping), with and without
        # late keys, in chunks small enough that a shard is several of them
        for name, keys, seed_rows in (("host_late", k[lo:hi], 4096), ("host_seeded", k[lo:hi] % (7919 * 700), 65_536)):
            rd.SEED_ROWS = seed_rows
            hi_, hf, hu, hs = rd.sharded_groupby_sum_host(keys, v[lo:hi], ops=ops, shard_sizes=[n // 3, n - n // 3], chunk_rows=50_000)
            out[name] = (hi_.copy(), hf, hu, hs)
        q.put((rank, out))
    finally:
        tdist.destroy_process_group()


def test_two_rank_nccl_paths():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 400_000
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k, f, v, s = _data(n)
    eik, eifk, eu = orc.MultiKeyGroupBy32([k], 0, None, 2)
    np.testing.assert_array_equal(np.concatenate([got[0]["ikey"], got[1]["ikey"]]), eik)
    cnt, ig, ifg = orc.BinCount(eik, eu + 1, pack=True)
    for r in range(2):
        assert got[r]["u"] == eu
        np.testing.assert_array_equal(got[r]["ifk"], eifk)
        for fn in (0, 1, 3, 5):
            exp = orc.GroupByAll32([v], eik, eu, [fn], [0], [eu + 2], 0)[0]
            np.testing.assert_allclose(got[r][f"r{fn}"], exp, rtol=1e-11, atol=1e-9, equal_nan=True)
        np.testing.assert_array_equal(got[r]["first"][1:], orc.GroupByAllPack32([v], eik, ig, ifg, cnt, eu, [100], [1], [eu + 1], 0)[0][1:])
        np.testing.assert_array_equal(got[r]["last"][1:], orc.GroupByAllPack32([v], eik, ig, ifg, cnt, eu, [102], [1], [eu + 1], 0)[0][1:])
    for r in range(2):
        pk = got[r]["pack"]
        np.testing.assert_array_equal(pk["nCountGroup"], cnt)
        np.testing.assert_array_equal(pk["iFirstGroup"], ifg)
        lo_p = int(ifg[pk["bin_lo"]]) if pk["bin_lo"] < len(ifg) else len(ig)
        hi_p = int(ifg[pk["bin_hi"]]) if pk["bin_hi"] < len(ifg) else len(ig)
        np.testing.assert_array_equal(pk["iGroup_owned"], ig[lo_p:hi_p])
    np.testing.assert_allclose(np.concatenate([got[0]["cumsum"], got[1]["cumsum"]]),
                               orc.EmaAll32([v], eik, eu, 300, (0.0, None, None, None))[0], rtol=1e-12, atol=1e-9)
    eik2, _, eu2 = orc.MultiKeyGroupBy32([k, s], 0, None, 2)
    np.testing.assert_array_equal(np.concatenate([got[0]["ikey2"], got[1]["ikey2"]]), eik2)
    np.testing.assert_allclose(got[0]["sum2"], orc.GroupByAll32([v], eik2, eu2, [0], [0], [eu2 + 2], 0)[0], rtol=1e-11, atol=1e-9)
    np.testing.assert_array_equal(np.concatenate([got[0]["sort"], got[1]["sort"]]), np.lexsort([f, s, k]))
    eik, eifk, ecnt = orc.GroupFromLexSort(orc.LexSort32([f, s]), [f, s])
    g0, g1 = got[0]["gfl"], got[1]["gfl"]
    np.testing.assert_array_equal(np.concatenate([g0["ikey"], g1["ikey"]]), eik)
    assert g0["unique_count"] == g1["unique_count"] == len(eifk)
    np.testing.assert_array_equal(g1["iFirstKey"], eifk)
    np.testing.assert_array_equal(g0["nCountGroup"], ecnt)
    np.testing.assert_array_equal(np.concatenate([g0["nCountGroup_owned"], g1["nCountGroup_owned"]]), ecnt[1:])
    for name, keys in (("host_late", k), ("host_seeded", k % (7919 * 700))):
        eik, eifk, eu = orc.MultiKeyGroupBy32([keys], 0, None, 2)
        exp = orc.GroupByAll32([v], eik, eu, [0], [1], [eu + 1], 0)[0]
        np.testing.assert_array_equal(np.concatenate([got[0][name][0], got[1][name][0]]), eik, err_msg=name)
        for r in range(2):
            gi, gf, gu, gs = got[r][name]
            assert gu == eu and gi.dtype == np.int32
            np.testing.assert_array_equal(gf, eifk, err_msg=name)
            np.testing.assert_allclose(gs[1:], exp[1:], rtol=1e-12, atol=1e-9, err_msg=name)
    for name, keys in (("gs_late", k), ("gs_seeded", k % (7919 * 700)), ("gs_seeded_small_prefix", k % (7919 * 700))):
        eik, eifk, eu = orc.MultiKeyGroupBy32([keys], 0, None, 2)
        for vname, vals in (("f64", v), ("f32", v.astype(np.float32)), ("i32", (v * 100).astype(np.int32))):
            exp = orc.GroupByAll32([vals], eik, eu, [0], [1], [eu + 1], 0)[0]
            np.testing.assert_array_equal(np.concatenate([got[0][f"{name}_{vname}"][0], got[1][f"{name}_{vname}"][0]]), eik, err_msg=name)
            for r in range(2):
                gi, gf, gu, gs = got[r][f"{name}_{vname}"]
                assert gu == eu
                np.testing.assert_array_equal(gf, eifk, err_msg=name)
                assert gs.dtype == exp.dtype
                if vname == "i32":
                    np.testing.assert_array_equal(gs[1:], exp[1:], err_msg=name)
                else:
                    np.testing.assert_allclose(gs[1:], exp[1:], rtol=1e-5 if vname == "f32" else 1e-12, atol=1e-3 if vname == "f32" else 1e-9, err_msg=name)
