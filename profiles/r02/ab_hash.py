"""A/B of the hash stage: RTB_LIB selects the library build; prints the steady-state round rates (RTB_GB_TRACE) and
CUDA-event times of rc.MultiKeyGroupBy32 / rc.GroupByAll32 on the C2 recipe."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
import riptable_b200.rc as rc

dev = torch.device("cuda", 0)
rows = int(os.environ.get("ROWS", 1_000_000_000))
key, val = bench.make_data(torch, dev, rows, 1_000_000, 0)
hint = int(os.environ.get("HINT", 0))
for it in range(4):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    ik, ifk, u = rc.MultiKeyGroupBy32([key], hint, None, 2)
    e1.record()
    s = rc.GroupByAll32([val], ik, u, [0], [1], [u + 1], 0)[0]
    e2.record()
    torch.cuda.synchronize()
    print(f"lib={os.environ.get('RTB_LIB','current')} hint={hint} it={it} hash {e0.elapsed_time(e1):.3f} ms sum {e1.elapsed_time(e2):.3f} ms U={u}", flush=True)
