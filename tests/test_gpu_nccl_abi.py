"""GPU, needs >= 2 devices: the multi-GPU C ABI (include/riptable_b200_nccl.h) driven with a raw ncclComm_t created through
libnccl itself (no torch.distributed): the sharded group-and-sum step against the oracle on the concatenated rows."""
import os

import numpy as np
import pytest

from oracle import riptide_oracle as orc


def test_nccl_library_exports_its_symbols():
    """CPU: libriptable_b200_nccl.so loads (against libriptable_b200.so and libnccl.so.2) and exports what the header declares."""
    import re

    from riptable_b200 import build, nccl_abi

    build.build_nccl()
    assert not nccl_abi.MISSING
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "riptable_b200_nccl.h")).read()
    declared = set(re.findall(r"^(?:size_t|int)\s+(rtb_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(nccl_abi.PROTOTYPES), declared ^ set(nccl_abi.PROTOTYPES)
    assert nccl_abi.lib.rtb_sharded_groupby_sum32_workspace_bytes(1_000_000, 100_000, 8) > 0


def _data(n, late):
    rng = np.random.default_rng(8)
    k = rng.integers(-(2**40), 2**40, 5000)[rng.integers(0, 5000, n)]
    if late:
        k[n - 1000:] = 2**50 + np.arange(1000)  # keys shard 0's prefix cannot contain
    v = rng.random(n) * 2 - 1
    return k, v


def _worker(rank, world, uid, n, late, vdt, q):
    import ctypes as C

    import torch

    torch.cuda.set_device(rank)
    from riptable_b200 import nccl_abi
    from riptable_b200._lib import rtb_column
    from riptable_b200.enums import dtype_code

    comm = nccl_abi.comm_init_rank(uid, world, rank)
    try:
        k, v = _data(n, late)
        v = v.astype(vdt) if np.dtype(vdt).kind == "f" else (v * 100).astype(vdt)
        cuts = [0, n // 3, n] if world == 2 else np.linspace(0, n, world + 1).astype(int).tolist()
        lo, hi = cuts[rank], cuts[rank + 1]
        kd = torch.from_numpy(k[lo:hi].copy()).cuda()
        vd = torch.from_numpy(v[lo:hi].copy()).cuda()
        m, mu = hi - lo, 1 << 16
        ikey = torch.empty(max(m, 4), dtype=torch.int32, device="cuda")
        ifk = torch.empty(mu, dtype=torch.int64, device="cuda")
        sums = torch.empty(mu + 1, dtype=torch.int64, device="cuda")
        ws = torch.empty(nccl_abi.lib.rtb_sharded_groupby_sum32_workspace_bytes(max(c1 - c0 for c0, c1 in zip(cuts, cuts[1:])), mu, 8),
                         dtype=torch.uint8, device="cuda")
        col = rtb_column(C.c_void_p(kd.data_ptr()), dtype_code(np.dtype(np.int64)), 8)
        res = nccl_abi.rtb_sharded_result()
        rc = nccl_abi.lib.rtb_sharded_groupby_sum32_nccl(
            comm, rank, world, C.byref(col), C.c_void_p(vd.data_ptr()), dtype_code(np.dtype(vdt)), 0, m, 131_072,
            C.c_void_p(ikey.data_ptr()), C.c_void_p(ifk.data_ptr()), C.c_void_p(sums.data_ptr()), mu, C.byref(res),
            C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        u = int(res.unique_count)
        raw = sums[: u + 1]
        q.put((rank, rc, u, int(res.late_keys), ikey[:m].cpu().numpy(), ifk[:u].cpu().numpy(),
               (raw.view(torch.float64) if np.dtype(vdt).kind == "f" else raw).cpu().numpy()))
    finally:
        nccl_abi.comm_destroy(comm)


@pytest.mark.gpu
@pytest.mark.parametrize("vdt", [np.float64, np.float32, np.int32])
def test_sharded_groupby_sum_over_a_raw_nccl_communicator(vdt):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from riptable_b200 import nccl_abi

    n = 600_000
    ctx = mp.get_context("spawn")
    for late in (False, True):
        uid = nccl_abi.get_unique_id()
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, uid, n, late, vdt, q)) for r in range(2)]
        for p in procs:
            p.start()
        got = {r[0]: r for r in (q.get(timeout=240) for _ in range(2))}
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        k, v = _data(n, late)
        v = v.astype(vdt) if np.dtype(vdt).kind == "f" else (v * 100).astype(vdt)
        ok, ofk, ou = orc.MultiKeyGroupBy32([k])
        if late:
            # keys outside the seed: every rank reports it (same positive status); the seeded part of iKey is already global
            assert got[0][1] == got[1][1] == nccl_abi.RTB_SHARDED_LATE_KEYS and got[0][3] == got[1][3] > 0
            continue
        assert got[0][1] == got[1][1] == 0 and got[0][2] == got[1][2] == ou
        np.testing.assert_array_equal(np.concatenate([got[0][4], got[1][4]]), ok)
        exp = orc.GroupByAll32([v], ok, ou, [0], [1], [ou + 1], 0)[0]
        for r in range(2):
            np.testing.assert_array_equal(got[r][5], ofk)
            if np.dtype(vdt).kind == "f":
                np.testing.assert_allclose(got[r][6][1:], exp[1:].astype(np.float64), rtol=1e-5 if vdt == np.float32 else 1e-12, atol=1e-3 if vdt == np.float32 else 1e-9)
            else:
                np.testing.assert_array_equal(got[r][6][1:], exp[1:])
