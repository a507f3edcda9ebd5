#!/usr/bin/env python
"""bench.py — groupby-sum rows/s on BASELINE.json's C2 workload (int64 key, 1M uniques, f64 value).

One "step" = one pass of the hot path over one batch of synthetic rows, producing what riptable's
    rc.MultiKeyGroupBy32([key])            -> iKey, iFirstKey, unique_count      (a1)
    rc.GroupByAll32([value], iKey, ... GB_SUM)                                   (a5)
produce: iKey[N], iFirstKey[U], U and sums[U+1].  The step is ONE call of rc.MultiKeyGroupBySum32 (the fused
group-and-sum entry, rtb_multikey_groupby_sum32); the same outputs through the two separate calls are timed beside it
(`two_call`).  For N > 1 the step is riptable_b200.dist.sharded_groupby_sum: seeded shard grouping with the sums fused
in, then one all-reduce of the per-bin partial sums.

  value : whole-job rows/s with inputs resident in HBM (device mode of the rc shim), CUDA-event timed.
  e2e   : same metric through the host-buffer (NumPy) mode of the rc shim: pinned host inputs,
          H2D of key+value and D2H of iKey/iFirstKey/sums inside the timed region.
  roofline : the dominant kernel's algorithmic bytes / its own CUDA-event duration (rtb_profile_*),
          against MEASURED_PEAKS.json's HBM copy bandwidth.
  cpu_baseline : the C restatement of the reference algorithm (oracle/oracle.c) on a fixed 1e8-row sample of the
          same recipe, which also checks the GPU results of that sample (`parity`).
  strong : (N > 1) the same step over 1 B rows TOTAL, N/G rows per rank (BASELINE.json configs[1]).
  parity_check : (N > 1) a 4 Mi-row-per-rank sharded run compared bit for bit (iKey, iFirstKey) and to 1e-12 (sums)
          with a single-GPU run over the concatenated rows.

`--impl reference` times that CPU restatement instead (riptide_cpp itself is not available: SURVEY 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "groupby-sum rows/s (int64 key, 1M uniques, f64 value)"
UNIT = "rows/s"
ALG_BYTES_HASH = 12   # int64 key read + int32 iKey write                         
ALG_BYTES_SUM = 12    # int32 iKey read + f64 value read                          
ALG_BYTES_FUSED = 20  # int64 key read + f64 value read + int32 iKey write, once each (the fused kernel's compulsory traffic)
CPU_SAMPLE_ROWS = 100_000_000  # the CPU leg runs N = 1e8, whatever --steps says


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=int(os.environ.get("RTB_BENCH_ROWS", 1_000_000_000)),
                    help="rows per GPU (weak scaling: every rank runs the C2 shard)")
    ap.add_argument("--uniques", type=int, default=1_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=CPU_SAMPLE_ROWS, help="rows of the CPU sample (fixed: independent of --steps)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (N > 1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity check (N > 1)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread every ~2 ms (an
    `nvidia-smi -lms` child needs about a second to start on an 8-GPU box, longer than the timed region), with the
    nvidia-smi loop as the fallback when the NVML binding is missing."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        self.index, self.uuid, self.proc, self.lines = index, uuid, None, []
        self.nvml = self.handle = None
        self.sm, self.mx, self.reason_bits, self.run = [], None, 0, False
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.handle = pynvml, h
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml:
            self.run = True
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml:
            self.run = False
            self.t.join(timeout=1)
            nv = self.nvml
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            reasons = sorted(k for k, b in bits.items() if self.reason_bits & b)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                    "samples": len(self.sm), "reasons": reasons, "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
def make_data(torch, dev, rows, uniques, seed):
    """C2 recipe: key = U_vals[idx], U_vals = `uniques` distinct int64 drawn from the full
    int64 range (non-consecutive), value f64 uniform(-1, 1).  Generated on the device, seeded."""
    g = torch.Generator(device=dev)
    g.manual_seed(12345)  # the unique key values are the same on every rank
    uvals = torch.randint(-(2**63), 2**63 - 1, (uniques,), dtype=torch.int64, device=dev, generator=g)
    g.manual_seed(12345 + 7919 * (seed + 1))
    key = torch.empty(rows, dtype=torch.int64, device=dev)
    val = torch.empty(rows, dtype=torch.float64, device=dev)
    chunk = 1 << 27
    for s in range(0, rows, chunk):
        e = min(rows, s + chunk)
        idx = torch.randint(0, uniques, (e - s,), dtype=torch.int64, device=dev, generator=g)
        torch.index_select(uvals, 0, idx, out=key[s:e])
        val[s:e].uniform_(-1.0, 1.0, generator=g)
        del idx
    return key, val


def cpu_arm(rows, uniques, steps, warmup, sample_rows=CPU_SAMPLE_ROWS, check_gpu=False):
    """The CPU restatement (oracle/oracle.c) on a fixed-size sample of the same recipe: `sample_rows` rows per step
    (1e8 by default) whatever the step count, so the figure does not depend on --steps.
    check_gpu: also run the GPU path on that very sample and compare (the oracle as the checker)."""
    from oracle import c_oracle

    c_oracle.build()
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core this process may run on
    threads = max(1, len(os.sched_getaffinity(0)))
    rng = np.random.default_rng(12345)
    uvals = rng.integers(-(2**63), 2**63 - 1, uniques, dtype=np.int64)
    n = int(min(rows, sample_rows)) if sample_rows > 0 else int(min(rows, CPU_SAMPLE_ROWS))
    key = uvals[rng.integers(0, uniques, n)]
    val = rng.random(n) * 2.0 - 1.0
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        ik, ifk, u = c_oracle.groupby_u64(key, threads=threads)
        s = c_oracle.sum_f64(val, ik, u, threads=threads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    total = sum(ts)
    out = {"value": n * len(ts) / total, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": f"{n} rows of the C2 recipe per step, {len(ts)} steps; hash stage on {threads} threads (riptable's "
                     f"cutoffs partitions + hstack_groupings merge), sum stage {threads} OpenMP threads over bin ranges",
           "ms_per_step": 1e3 * total / len(ts), "rows": n}
    if check_gpu:
        try:
            import riptable_b200.rc as rc

            gik, gifk, gu, gs = rc.MultiKeyGroupBySum32([key], val)
            ok = gu == u and np.array_equal(gik, ik) and np.array_equal(gifk, ifk)
            err = float(np.max(np.abs(gs[1:] - s[1:]) / np.maximum(1.0, np.abs(s[1:])))) if u else 0.0
            out["parity"] = ("ok" if ok and err <= 1e-12 else "FAILED") + f": GPU fused call vs the CPU port on this sample, iKey / iFirstKey bit-exact = {ok}, max rel sum error {err:.2e}"
        except Exception as ex:
            out["parity"] = f"not run: {ex!r}"
    return out


def pin_to_gpu_cpus(local):
    """One rank per GPU on one host: run this rank (and first-touch its pinned host buffers) on the CPUs next to its GPU,
    what `numactl --cpunodebind` does for a multi-process launch.  Best effort: None when NVML cannot say."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------------
def main():
    a = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))

    if a.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_arm(a.rows, a.uniques, max(a.steps, 1), max(a.warmup, 1), a.cpu_rows)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int64 keys / f64 sums", "data": "synthetic",
                "config": {"workload": "C2: int64 key (1M uniques drawn from the full int64 range), gb sum of f64",
                           "rows_per_step": cb["rows"], "uniques": a.uniques,
                           "sample": "fixed 1e8-row sample per step (BASELINE.md section 3), independent of --steps"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "rows")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_cpus(local) if world > 1 else None
    import riptable_b200.rc as rc
    from riptable_b200 import _lib
    from riptable_b200.enums import GB_FUNCTIONS

    dist = None
    force_dist = world == 1 and os.environ.get("RTB_BENCH_FORCE_DIST") == "1"  # exercise the shard-merge path on one GPU
    if force_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
    if world > 1 or force_dist:
        import torch.distributed as tdist

        tdist.init_process_group("nccl", device_id=dev, rank=rank, world_size=world)
        from riptable_b200 import dist as rdist

        dist = rdist

    rows, U = a.rows, a.uniques
    key, val = make_data(torch, dev, rows, U, rank)
    torch.cuda.synchronize()

    def step_device(k=None, v=None):
        k = key if k is None else k
        v = val if v is None else v
        if dist is None:
            return rc.MultiKeyGroupBySum32([k], v, 0, None, 2, GB_FUNCTIONS.GB_SUM, 1)
        return dist.sharded_groupby_sum([k], v, shard_sizes=[int(k.numel())] * world)  # equal shards: sizes known

    def step_two_call():
        ik, ifk, u = rc.MultiKeyGroupBy32([key], 0, None, 2)
        s = rc.GroupByAll32([val], ik, u, [GB_FUNCTIONS.GB_SUM], [1], [u + 1], 0)[0]
        return ik, ifk, u, s

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize on both sides, CUDA events, max over ranks -> ms."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        t_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([t_ms], dtype=torch.float64, device=dev)
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms, out

    for _ in range(max(a.warmup, 3)):
        out = step_device()
    del out
    barrier()
    _lib.lib.rtb_profile_enable(1)
    _lib.lib.rtb_profile_reset()
    launches0 = _lib.launch_count()
    try:
        dev_uuid = torch.cuda.get_device_properties(local).uuid
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local, dev_uuid)
    sampler.start()
    ms, out = timed(step_device, a.steps)
    clocks = sampler.stop()
    launches = _lib.launch_count() - launches0
    uniq_found = int(out[2])
    del out
    prof = {name: _lib.profile_get(pid) for name, pid in (("gb_round", 0), ("accum", 1))}
    _lib.lib.rtb_profile_enable(0)
    value = rows * world * a.steps / (ms / 1e3)

    # ---- the same outputs through riptable's two separate calls (N = 1), for comparison with round 1 ------------
    two_call = None
    if dist is None:
        for _ in range(2):
            out = step_two_call()
        del out
        _lib.lib.rtb_profile_enable(1)
        _lib.lib.rtb_profile_reset()
        ms2, out = timed(step_two_call, a.steps)
        del out
        p2 = {name: _lib.profile_get(pid) for name, pid in (("gb_round", 0), ("accum", 1))}
        _lib.lib.rtb_profile_enable(0)
        two_call = {"value": rows * a.steps / (ms2 / 1e3), "unit": UNIT, "ms_per_step": ms2 / a.steps,
                    "hash_ms_per_step": p2["gb_round"][0] / a.steps, "sum_ms_per_step": p2["accum"][0] / a.steps,
                    "note": "rc.MultiKeyGroupBy32 then rc.GroupByAll32(GB_SUM): two sweeps, 24 B/row algorithmic"}

    # ---- strong scaling (N > 1): 1 B rows in total, rows / N per rank (BASELINE.json configs[1]) ----------------
    strong = None
    if dist is not None and world > 1 and not a.no_strong:
        srows = (rows // world) & ~1023
        sk, sv = key[:srows], val[:srows]
        for _ in range(3):
            out = step_device(sk, sv)
        del out
        ms_s, out = timed(lambda: step_device(sk, sv), a.steps)
        del out
        strong = {"rows_total": srows * world, "rows_per_gpu": srows, "ms_per_step": ms_s / a.steps,
                  "value": srows * world * a.steps / (ms_s / 1e3), "unit": UNIT, "scaling": "strong",
                  "note": "same step, the C2 row count divided across the ranks; efficiency = value / (N x the N=1 line's value)"}

    # ---- multi-GPU parity (N > 1): sharded run == single-GPU run over the concatenated rows ----------------------
    parity = None
    if dist is not None and not a.no_parity:
        try:
            prow = min(rows, 1 << 22)
            pk, pv = key[:prow].contiguous(), val[:prow].contiguous()
            gik, gifk, gu, gs = dist.sharded_groupby_sum([pk], pv)  # (this call also exercises the shard-size all-gather)
            allk = torch.empty(prow * world, dtype=pk.dtype, device=dev)
            allv = torch.empty(prow * world, dtype=pv.dtype, device=dev)
            alli = torch.empty(prow * world, dtype=torch.int32, device=dev)
            tdist.all_gather_into_tensor(allk, pk)
            tdist.all_gather_into_tensor(allv, pv)
            tdist.all_gather_into_tensor(alli, gik.contiguous())
            sik, sifk, su = rc.MultiKeyGroupBy32([allk], 0, None, 2)
            ss = rc.GroupByAll32([allv], sik, su, [GB_FUNCTIONS.GB_SUM], [1], [su + 1], 0)[0]
            ok_i = bool(su == gu and torch.equal(sik, alli) and torch.equal(sifk.to(torch.int64), gifk.to(torch.int64)))
            err = float(((ss[1:] - gs[1:]).abs() / ss[1:].abs().clamp(min=1.0)).max()) if su else 0.0
            flag = torch.tensor([1 if (ok_i and err <= 1e-12) else 0], dtype=torch.int32, device=dev)
            tdist.all_reduce(flag, op=tdist.ReduceOp.MIN)
            parity = {"status": "ok" if int(flag.item()) == 1 else "FAILED", "rows_per_rank": prow, "unique_count": int(su),
                      "ikey_ifirstkey_bit_exact": ok_i, "max_rel_sum_error": err,
                      "against": "single-GPU rc.MultiKeyGroupBy32 + rc.GroupByAll32 over the all-gathered rows of every rank"}
            del allk, allv, alli, sik, ss
        except Exception as ex:
            parity = {"status": f"not run: {ex!r}"}
            try:
                barrier()
            except Exception:
                pass

    # ---- end to end through the host-buffer API ----------------------------------------------------------------
    # N = 1: the rc shim in host mode (NumPy in, NumPy out).  N > 1: every rank uploads its shard from pinned host
    # memory, runs the sharded step, and reads iKey / iFirstKey / sums back to the host.  With several ranks on one
    # host the shard is bounded (pinned memory and host DRAM bandwidth are shared by all ranks).
    e2e = None
    if not a.no_e2e:
        try:
            erows = rows if world == 1 else min(rows, 250_000_000)
            hk = torch.empty(erows, dtype=torch.int64, pin_memory=True)
            hv = torch.empty(erows, dtype=torch.float64, pin_memory=True)
            hk.copy_(key[:erows]); hv.copy_(val[:erows])
            torch.cuda.synchronize()
            nk, nv = hk.numpy(), hv.numpy()
            del key, val
            torch.cuda.empty_cache()

            if dist is None:
                def step_host():
                    return rc.MultiKeyGroupBySum32([nk], nv, 0, None, 2, GB_FUNCTIONS.GB_SUM, 1)
                h2d = erows * 8 + erows * 8  # key and value; the fused call never hands iKey back to the device
            else:
                def step_host():
                    return dist.sharded_groupby_sum_host(nk, nv, shard_sizes=[erows] * world)
                h2d = erows * 8 + erows * 8

            e2e_steps = max(1, min(a.steps, 3))
            # warm-up: two host steps with the first result still alive, so that torch's pinned-memory cache holds both
            # result buffers the timed loop alternates between (a first cudaHostAlloc of 4 GB takes seconds)
            keep = step_host()
            keep2 = step_host()
            del keep, keep2
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                ik, ifk, u, s = step_host()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
                dt = float(t.item())
            d2h = erows * 4 + u * (4 if dist is None else 8) + (u + 1) * 8
            e2e = {"value": erows * world * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "steps": e2e_steps, "rows_per_gpu": erows, "cpus_bound_per_rank": numa,
                   "note": ("rc.MultiKeyGroupBySum32 in host mode: pinned NumPy inputs in, NumPy iKey/iFirstKey/sums out; PCIe-bound" if dist is None else
                            "per rank: dist.sharded_groupby_sum_host -- the pinned host shard streams to the device in 32 Mi-row "
                            "chunks under the sharded groupby-sum, iKey chunks stream back, iFirstKey/sums at the end; "
                            "PCIe-bound; bytes are per rank")}
            del hk, hv, nk, nv, ik, s
        except Exception as ex:  # an end-to-end leg that cannot run (e.g. pinned memory) must not take the bench line down
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": f"failed: {ex!r}"}
            if world > 1:
                try:
                    barrier()
                except Exception:
                    pass

    if rank != 0:
        if world > 1:
            tdist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    kern = {}
    fused_rows = prof["gb_round"][2] - prof["accum"][2]  # rows of the fused rounds (unfused rounds are re-read by accum)
    for name, (pms, pl, pu) in prof.items():
        if pl:
            # gb_round: key in + iKey out for every row, + the value stream for the rows of the fused rounds
            nbytes = pu * ALG_BYTES_HASH + max(fused_rows, 0) * 8 if name == "gb_round" else pu * ALG_BYTES_SUM
            kern[name] = {"ms_total": pms, "launches": pl, "rows": pu, "alg_bytes_per_row": nbytes / max(pu, 1),
                          "achieved_gbs": nbytes / (pms / 1e3) / 1e9}
    dom = max(kern, key=lambda k: kern[k]["ms_total"]) if kern else None
    roofline = None
    if dom:
        k = kern[dom]
        roofline = {"bound": "hbm", "kernel": "gb_round_direct<8,0,1> (fused probe + sum; + gb_round_insert<8> in the discovery rounds)" if dom == "gb_round" else "accum_kernel<OP_SUMF,false>",
                    "achieved": k["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": k["achieved_gbs"] / peak,
                    "traffic": None, "peak_source": peak_src,
                    "alg_bytes_per_launch": k["rows"] * k["alg_bytes_per_row"] / k["launches"],
                    "avg_launch_ms": k["ms_total"] / k["launches"]}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                per_row = json.load(open(tp))[dom + "_fused" if dom == "gb_round" else dom]["dram_bytes_per_row"]  # committed ncu --set full capture
                roofline["traffic"] = per_row * k["rows"] / k["launches"]
                roofline["traffic_source"] = "profiles/traffic.json (ncu dram bytes per row x rows per launch)"
            except Exception:
                pass
    step_bytes = ALG_BYTES_FUSED * rows
    cb = None
    if world > 1:
        cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "timed at N=1 only (see --impl reference)"}
    elif not a.no_cpu:  # the CPU leg is an N=1 item (rank 0 alone); --impl reference times it at any N
        try:
            cb = cpu_arm(rows, U, 2, 1, a.cpu_rows, check_gpu=True)
            cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "rows", "parity") if k in cb}
        except Exception as ex:  # the CPU leg must never take the GPU number down with it
            cb = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64 keys / f64 sums", "data": "synthetic",
        "config": {"workload": "C2: int64 key (1M uniques drawn from the full int64 range), gb sum of f64",
                   "step": ("one rc.MultiKeyGroupBySum32 call (fused group-and-sum): iKey, iFirstKey, U, sums" if world == 1 and not force_dist
                            else "dist.sharded_groupby_sum: seeded shard grouping with fused sums + all-reduce of the partials"),
                   "rows_per_gpu": rows, "uniques": U, "unique_count_found": uniq_found,
                   "e2e_rows_per_gpu": (e2e or {}).get("rows_per_gpu"),
                   "l2": "inputs (16 B/row x rows) are far larger than the 126 MB L2; no explicit flush",
                   "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline,
        "step_hbm_frac": step_bytes / (ms / a.steps / 1e3) / 1e9 / peak,
        "step_hbm_frac_24B": (ALG_BYTES_HASH + ALG_BYTES_SUM) * rows / (ms / a.steps / 1e3) / 1e9 / peak,
        "step_alg_bytes_per_row": {"fused": ALG_BYTES_FUSED, "two_call (SURVEY 8d)": ALG_BYTES_HASH + ALG_BYTES_SUM},
        "kernels": kern, "two_call": two_call, "strong": strong, "parity_check": parity,
        "cpu_baseline": cb,
    }
    print(json.dumps(line))
    if world > 1 or force_dist:
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
