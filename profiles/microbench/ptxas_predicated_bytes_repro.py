import numpy as np, sys, torch, ctypes as C, os
sys.path.insert(0, '.')
import riptable_b200.rc as rc
from riptable_b200._lib import lib, rtb_column
rng = np.random.default_rng(4)
n = 600
s = rng.choice(np.array([b"aa", b"b", b"", b"cccc"], dtype="S4"), n)
codes = np.zeros(n, np.uint64)
raw = s.view(np.uint8).reshape(n, 4)
for j in range(4):
    codes |= raw[:, j].astype(np.uint64) << np.uint64(56 - 8 * j)
exp = np.lexsort([s])
res = []
for rep in range(8):
    d = rc._to_dev(s)
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    col = rc._col(d)
    lib.rtb_debug_sort_codes.restype = C.c_int
    lib.rtb_debug_sort_codes.argtypes = [C.POINTER(rtb_column), C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.rtb_debug_sort_codes(C.byref(col), n, 0, 0, C.c_void_p(out.data_ptr()), None)
    got = out.cpu().numpy().view(np.uint64)
    res.append(int((got != codes).sum()))
    a = rc.LexSort32([s])
    res.append(int((a != exp).sum()))
print("sync" if os.environ.get("RTB_DEBUG_SYNC") else "async", "(codes bad, sort bad) x8:", res)
bad = np.flatnonzero(got != codes)
print("bad rows", bad[:40])
for r in bad[:12]:
    print(r, s[r], hex(int(got[r])), hex(int(codes[r])))
