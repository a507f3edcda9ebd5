import sys, torch
sys.path.insert(0, '.')
import riptable_b200.rc as rc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000_000
g = torch.Generator(device='cuda'); g.manual_seed(1)
c = torch.randint(0, 1_000_000, (n,), device='cuda', generator=g, dtype=torch.int32)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); p = rc.LexSort32([c]); e1.record(); torch.cuda.synchronize()
    print("lexsort int32 1e6 values", n, "rows", e0.elapsed_time(e1), "ms")
