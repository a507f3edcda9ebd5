import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import riptable_b200.rc as rc
from oracle import riptide_oracle as orc
rng = np.random.default_rng(5)
n = 6_000_000
k = rng.integers(0, 2000, n).astype(np.int64) * 7919
late = rng.integers(4_500_000, n, 3000)
k[late] = 10**12 + rng.integers(0, 500, len(late))
v = rng.random(n) * 2 - 1
ok, ofk, ou = orc.MultiKeyGroupBy32([k])
ik0, ifk0, u0 = rc.MultiKeyGroupBy32([k], 0, None, 2)
print("unfused equal:", np.array_equal(ik0, ok), u0, ou)
ik, ifk, u, s = rc.MultiKeyGroupBySum32([k], v)
bad = np.flatnonzero(ik != ok)
print("fused mismatches:", len(bad), "u", u, ou, "ifk equal", np.array_equal(ifk, ofk))
print("rows", bad[:10], "got", ik[bad[:10]], "exp", ok[bad[:10]])
print("got range", ik[bad].min() if len(bad) else None, ik[bad].max() if len(bad) else None)
exp = orc.GroupByAll32([v], ok, ou, [0], [1], [ou + 1], 0)[0]
print("sum err", np.abs(s[1:] - exp[1:]).max())
