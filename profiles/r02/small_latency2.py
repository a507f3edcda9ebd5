"""Where a small rc.MultiKeyGroupBy32 call spends its time: the whole Python shim against the C-ABI call alone (buffers
allocated once), per input size.  RTB_GB_SMALL=0: round schedule, 1: single cooperative launch."""
import ctypes as C, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import riptable_b200.rc as rc
from riptable_b200._lib import lib, rtb_column

out = {}
for n in (10_000, 100_000, 1_000_000):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    k = torch.randint(0, 10_000, (n,), device="cuda", generator=g, dtype=torch.int32)
    for _ in range(5):
        rc.MultiKeyGroupBy32([k], 0, None, 2)
    torch.cuda.synchronize()
    reps = 300
    t0 = time.perf_counter()
    for _ in range(reps):
        rc.MultiKeyGroupBy32([k], 0, None, 2)
    torch.cuda.synchronize()
    shim = (time.perf_counter() - t0) / reps * 1e6
    mu = n
    ikey = torch.empty(n, dtype=torch.int32, device="cuda")
    ifk = torch.empty(mu, dtype=torch.int32, device="cuda")
    ws = torch.empty(lib.rtb_groupby_workspace_bytes_rows(n, mu, 0, 0), dtype=torch.uint8, device="cuda")
    col = (rtb_column * 1)(); col[0] = rtb_column(C.c_void_p(k.data_ptr()), 5, 4)
    u, need = C.c_int64(0), C.c_int64(0)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        return lib.rtb_multikey_groupby32(col, 1, n, 0, None, None, 0, C.c_void_p(ikey.data_ptr()), C.c_void_p(ifk.data_ptr()), mu,
                                          C.byref(u), None, C.byref(need), C.c_void_p(ws.data_ptr()), ws.numel(), stream)
    for _ in range(5):
        assert call() == 0
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    cabi = (time.perf_counter() - t0) / reps * 1e6
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record(); torch.cuda.synchronize()
    out[f"n{n}"] = {"shim_us": round(shim, 1), "c_abi_us": round(cabi, 1), "c_abi_device_us": round(e0.elapsed_time(e1) / reps * 1e3, 1), "U": u.value}
print(json.dumps({"small_path": os.environ.get("RTB_GB_SMALL", "1"), "per_call": out}))
