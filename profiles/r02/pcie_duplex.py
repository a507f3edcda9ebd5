"""Can host->device and device->host copies run concurrently at full rate on this box?  Pinned buffers, two streams."""
import json, time, torch
n = 1 << 30  # 1 GiB per buffer
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        if mode == "serial":
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    nbytes = reps * n * (2 if mode in ("both", "serial") else 1)
    return round(nbytes / dt / 1e9, 1)
run("both", 1)
print(json.dumps({m: run(m) for m in ("h2d", "d2h", "serial", "both")}))
