"""GPU: riptable_b200.dist through the product backend (CudaOps) on a single-rank NCCL group, against the
single-process oracle.  The multi-rank collectives are covered on CPU by tests/test_dist_gloo.py and on
2/4/8 GPUs by bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import riptide_oracle as orc


@pytest.fixture(scope="module")
def group():
    import torch
    import torch.distributed as tdist

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    tdist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield tdist.group.WORLD
    tdist.destroy_process_group()


@pytest.mark.parametrize("a2a", [1 << 40, 1])
def test_sharded_groupby_reduce_single_rank(group, a2a):
    import torch
    from riptable_b200 import dist as rd

    rng = np.random.default_rng(3)
    n = 300_000
    k1 = rng.integers(0, 5000, n).astype(np.int64) * 7919
    k2 = rng.choice(np.array([b"aa", b"b", b"", b"cccc"], dtype="S4"), n)
    v = rng.random(n) * 100 - 50
    v[rng.random(n) < 0.1] = np.nan
    vi = rng.integers(-1000, 1000, n).astype(np.int32)
    f = rng.random(n) < 0.8
    ops = rd.CudaOps()
    for keys, filt in (([k1], None), ([k1, k2], f)):
        g = rd.sharded_groupby(ops, [torch.from_numpy(k).cuda() if k.dtype.kind != "S" else k for k in keys],
                               None if filt is None else torch.from_numpy(filt).cuda(), group)
        eik, eifk, eu = orc.MultiKeyGroupBy32(keys, 0, filt, 2)
        assert g.unique_count == eu
        np.testing.assert_array_equal(g.ikey.cpu().numpy(), eik)
        np.testing.assert_array_equal(g.ifirstkey.cpu().numpy(), eifk)
        np.testing.assert_array_equal(rd.sharded_count(ops, g, group, a2a).cpu().numpy(), orc.BinCount(eik, eu + 1))
        vt, vit = torch.from_numpy(v).cuda(), torch.from_numpy(vi).cuda()
        for fn in (0, 1, 2, 3, 4, 5, 50, 51, 52, 53, 54, 55):
            got = rd.sharded_reduce(ops, vt, g, fn, group, a2a).cpu().numpy()
            exp = orc.GroupByAll32([v], eik, eu, [fn], [0], [eu + 2], 0)[0]
            np.testing.assert_allclose(got, exp, rtol=1e-11, atol=1e-9, equal_nan=True, err_msg=f"func {fn}")
        for fn in (0, 2, 3):
            got = rd.sharded_reduce(ops, vit, g, fn, group, a2a).cpu().numpy()
            np.testing.assert_array_equal(got, orc.GroupByAll32([vi], eik, eu, [fn], [0], [eu + 2], 0)[0])
    # first / nth / last and cumsum through the sharded entry points
    g = rd.sharded_groupby(ops, [torch.from_numpy(k1).cuda()], None, group)
    eik, eifk, eu = orc.MultiKeyGroupBy32([k1], 0, None, 2)
    cnt, ig, ifg = orc.BinCount(eik, eu + 1, pack=True)
    for fn, p in ((100, 0), (102, 0), (101, 2), (101, -1), (101, 100_000)):
        got = rd.sharded_pick(ops, torch.from_numpy(v).cuda(), g, fn, p, group).numpy()
        exp = orc.GroupByAllPack32([v], eik, ig, ifg, cnt, eu, [fn], [1], [eu + 1], p)[0]
        np.testing.assert_array_equal(got[1:].view(np.uint64), exp[1:].view(np.uint64))
        got = rd.sharded_pick(ops, k2, g, fn, p, group).numpy()
        np.testing.assert_array_equal(got[1:], orc.GroupByAllPack32([k2], eik, ig, ifg, cnt, eu, [fn], [1], [eu + 1], p)[0][1:])
    got = rd.sharded_cumsum(ops, vit, g, torch.from_numpy(f).cuda(), group).cpu().numpy()
    np.testing.assert_array_equal(got, orc.EmaAll32([vi], eik, eu, 300, (0.0, None, f, None))[0])
    got = rd.sharded_cumsum(ops, torch.from_numpy(np.nan_to_num(v)).cuda(), g, None, group).cpu().numpy()
    np.testing.assert_allclose(got, orc.EmaAll32([np.nan_to_num(v)], eik, eu, 300, (0.0, None, None, None))[0], rtol=1e-12, atol=1e-9)
    gl = rd.sharded_group_from_lexsort(ops, [torch.from_numpy(k1).cuda(), k2], group=group)
    eik, eifk, ecnt = orc.GroupFromLexSort(orc.LexSort32([k1, k2]), [k1, k2])
    np.testing.assert_array_equal(gl["ikey"].cpu().numpy(), eik)
    np.testing.assert_array_equal(gl["iFirstKey"].cpu().numpy(), eifk)
    np.testing.assert_array_equal(gl["nCountGroup"].cpu().numpy(), ecnt)
    ik, ifk, u, s = rd.sharded_groupby_sum([torch.from_numpy(k1).cuda()], torch.from_numpy(v).cuda(), group, ops)
    assert u == len(np.unique(k1)) and s.shape[0] == u + 1


@pytest.mark.parametrize("seed_rows,chunk_rows", [(4096, 50_000), (1 << 25, 100_000), (70_000, 1 << 25)])
def test_sharded_groupby_sum_host_single_rank(group, seed_rows, chunk_rows):
    """The host-shard form (HostFeed: the shard streams to the device chunk by chunk under the grouping, iKey chunks
    stream back) on one rank: late keys after a short seed prefix, a prefix that covers the shard, one chunk / many."""
    import torch
    from riptable_b200 import dist as rd

    rng = np.random.default_rng(12)
    n = 400_003
    k = rng.integers(0, 3000, n).astype(np.int64) * 7919
    k[: n // 2] %= 7919 * 1500  # half of the keys first appear in the second half
    v = rng.random(n) * 10 - 5
    old = rd.SEED_ROWS
    rd.SEED_ROWS = seed_rows
    try:
        for pinned, vals in ((True, v), (False, v.astype(np.float32)), (True, (v * 100).astype(np.int32))):
            hk = torch.from_numpy(k).pin_memory().numpy() if pinned else k  # pageable inputs go through the staging ring
            hv = torch.from_numpy(vals).pin_memory().numpy() if pinned else vals
            ik, ifk, u, s = rd.sharded_groupby_sum_host(hk, hv, group, chunk_rows=chunk_rows)
            eik, eifk, eu = orc.MultiKeyGroupBy32([k], 0, None, 2)
            exp = orc.GroupByAll32([vals], eik, eu, [0], [1], [eu + 1], 0)[0]
            assert u == eu and isinstance(ik, np.ndarray) and ik.dtype == np.int32
            np.testing.assert_array_equal(ik, eik)
            np.testing.assert_array_equal(ifk, eifk)
            assert s.dtype == exp.dtype
            if exp.dtype.kind == "f":
                np.testing.assert_allclose(s[1:], exp[1:], rtol=1e-5 if exp.dtype == np.float32 else 1e-12, atol=1e-3 if exp.dtype == np.float32 else 1e-9)
            else:
                np.testing.assert_array_equal(s[1:], exp[1:])
    finally:
        rd.SEED_ROWS = old
