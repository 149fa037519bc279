#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the radix-sort hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload keys|pairs]

One "step" = one complete sort of one batch of synthetic keys through vkrs_multi_sort (keys only: the
schedule vkrs_multi_sort picks for that N -- at 10^8 keys the bucket schedule: two unstable top-digit
partition passes + shared-memory sort of the 16-bit-prefix buckets; pairs: four stable LSD digit passes).
Workload at every N: BASELINE.json configs[1], 10^8 uniform random uint32 keys PER GPU
("scaling": "weak").  At N=1 that is exactly the configuration the metric is quoted on; at N>1
the N x 10^8 keys are sorted GLOBALLY by the bucket exchange of vkradixsort_b200/dist.py
(splitters -> stable partition -> all-to-all over NVLink -> local sort), rank r ending up with the
r-th key range.

Printed by rank 0, ONE JSON line:
  value      Mkeys/s, keys resident in HBM when the timed region starts (CUDA events on the launch
             stream around each sort, summed over K steps, max over ranks)
  e2e        the same sort through the host-buffer C-ABI call vkrs_multi_sort_host (pinned host
             keys -> H2D -> sort -> D2H), both copies inside the timed region
  roofline   the kernel with the largest share of the step: its algorithmic bytes per launch (8 B/key
             for a scatter or the local sort: one 4 B read + one 4 B write; 4 B/key for a histogram)
             over its CUDA-event duration, against MEASURED_PEAKS.json's copy bandwidth; every other
             kernel of the step the same way under roofline.kernels; plus the whole sort under
             BASELINE.json's 64 B/key formula
  cpu_baseline  single-thread std::sort of the same 10^8 keys on this box (the reference's own CPU
             arm, MultiRadixSort.cpp:141-146), rank 0, N=1 only

--impl reference times the CPU restatement of the reference's shaders (oracle/, OpenMP over work
groups, all host threads) on a bounded sample; the reference itself (Vulkan + glslc) cannot be
built in this image.  That leg and cpu_baseline are the only places this file touches oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_KEYS = 100_000_000          # BASELINE.json configs[1]
ALGO_BYTES_PER_KEY_SORT = 64  # BASELINE.json: 16 B/key/pass x 4 passes (keys only); 96 for pairs
ALGO_BYTES_PER_KEY_PASS = 8   # a scatter kernel / the local sort: one 4 B read + one 4 B write per key
ALGO_BYTES_PER_KEY_HIST = 4   # a histogram kernel: one 4 B read per key


def algo_bytes_per_key(kernel_name: str, pairs: bool) -> int:
    """Algorithmic HBM bytes per key and launch of one of the library's kernels."""
    if "histogram" in kernel_name:
        return ALGO_BYTES_PER_KEY_HIST
    if "scatter" in kernel_name or "local_tile" in kernel_name or "onesweep" in kernel_name:
        return ALGO_BYTES_PER_KEY_PASS * (2 if pairs else 1)
    return 0  # planning kernels: a few KB
METRIC = "Mkeys/s on 10^8 uint32 (1/2/4/8xB200); achieved HBM GB/s vs peak"
SEED = 0x5EED0002


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.005):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._dev = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._dev, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # no NVML: record that, never fail the bench for it
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            return self
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def _run(self):
        nv = self._nvml
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._dev, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._dev)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._dev)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        sm = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": sm, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def common_config(pairs: bool, world: int, n: int) -> dict:
    """The part of `config` both arms report identically."""
    return {"workload": ("10^8 uint32 key + uint32 payload pairs" if pairs else "10^8 random uint32 keys")
            + ", multi_radixsort (BASELINE.json configs[%d])" % (2 if pairs else 1)
            + (", per GPU; global sort by bucket exchange" if world > 1 else ""),
            "keys_per_gpu": n, "distribution": "uniform full-range uint32 (numpy PCG64)", "seed": SEED}


def make_keys(n: int, seed: int) -> np.ndarray:
    """Uniform full-range uint32 (BASELINE.json 'random uint32'); numpy PCG64, fixed seed."""
    return np.random.default_rng(seed).integers(0, 1 << 32, size=n, dtype=np.uint32)


# --------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of the reference's own shaders, all host threads
# --------------------------------------------------------------------------------------------
def run_reference(args) -> int:
    """The reference's algorithm on the host cores: ALL keys of the step (the same 10^8 keys, same seed, as the
    B200 arm), all host threads.  The thread count is set explicitly -- torchrun exports OMP_NUM_THREADS=1."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O  # the one other place bench.py may execute oracle/

    O.build()
    threads = os.cpu_count() or 1
    O.set_num_threads(threads)
    threads = O.num_threads()
    n = args.n
    nb = 4096            # the reference's published best nb at 10^8 keys (README timings: 43.49 ms on its GPU)
    pristine = make_keys(n, SEED)
    steps, warmup = args.steps, args.warmup
    t_probe0 = time.perf_counter()
    out = O.multi_sort(pristine, nb)[0]  # first (untimed) step: also tells how long a step takes on this host
    t_probe = time.perf_counter() - t_probe0
    assert np.all(out[1:] >= out[:-1])
    # keep the whole run within a few minutes whatever the host: cap the number of timed steps, never the workload
    budget_s = 150.0
    max_steps = max(1, int(budget_s / max(t_probe, 1e-3)) - 1)
    warm_done = 1
    while warm_done < warmup and warm_done + steps < max_steps:
        O.multi_sort(pristine, nb)
        warm_done += 1
    timed = max(1, min(steps, max_steps - warm_done))
    t_total = 0.0
    for _ in range(timed):
        t0 = time.perf_counter()
        out = O.multi_sort(pristine, nb)[0]  # copies the input (untimed part is small: one 400 MB memcpy), sorts in buf0/buf1
        t_total += time.perf_counter() - t0
    assert np.all(out[1:] >= out[:-1]) and int(out.astype(np.uint64).sum()) == int(pristine.astype(np.uint64).sum())
    ms = 1e3 * t_total / timed
    mkeys = n / (ms * 1e-3) / 1e6
    sample = min(n, 10_000_000)
    k1 = pristine[:sample].copy()
    std_ms = O.std_sort(k1)
    k2 = pristine.copy()
    par_ms = O.parallel_sort(k2)
    line = {
        "impl": "reference", "metric": METRIC, "value": mkeys, "unit": "Mkeys/s", "n_gpus": args.gpus,
        "steps": timed, "warmup": warm_done, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {**common_config(False, 1, n), "sample_keys_per_step": n, "steps_requested": steps,
                   "threads": threads, "nb": nb},
        "cpu_baseline": {"value": mkeys, "unit": "Mkeys/s", "cores": threads, "kind": "port",
                         "sample": f"all {n} keys of the step, restated multi_radixsort shaders "
                                   f"(oracle/vkrs_oracle.c, checked against the reference's own shaders in oracle/_ref; nb={nb}), "
                                   f"OpenMP over work groups, {threads} threads set explicitly",
                         "also_mkeys_per_s": {f"std_sort_1_thread_{sample}_keys": sample / std_ms / 1e3,
                                              f"gnu_parallel_sort_{threads}_threads_{n}_keys": n / par_ms / 1e3}},
        "e2e": {"value": mkeys, "unit": "Mkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the Vulkan host program of the reference cannot be built in this image (no Vulkan headers/loader, no glslc); "
                "this is its algorithm restated in C on the host cores (the restatement is pinned to the reference's shader "
                "source by tests/test_ref_shaders.py)",
    }
    print(json.dumps(line), flush=True)
    return 0



# --------------------------------------------------------------------------------------------
# BASELINE.json configs[2..4] (+ the 64-bit variant and the literal LSD loop): one record each
# --------------------------------------------------------------------------------------------
def other_configs(torch, handle, capi, dev, rank, world, dist, sorter):
    """{name: {"ms", "mkeys_s", "verified", ...}}: few repetitions each, CUDA events around the call, inputs restored
    before every repetition.  N = 1: pairs_1e8 (configs[2]), u64_5e7 (the reference's SORT_TYPE uint64_t variant),
    lsd_1e8 (the literal MultiRadixSort::execute loop), nsweep (configs[3], 28-bit keys as the reference's sweep),
    keys_8e8 (configs[4] at one GPU).  N > 1: strong_8e8 (configs[4] as stated: 8 * 10^8 keys in total)."""
    FLIP = -(1 << 31)
    out = {}

    def timed(fn, restore, reps=3, warm=1):
        for _ in range(warm):
            restore()
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            restore()
            if dist is not None:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, r

    def gen(n, seed, bits=32):
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        k = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
        return k if bits == 32 else k & ((1 << bits) - 1)

    def is_sorted(t):
        o = t ^ FLIP
        return bool((o[1:] >= o[:-1]).all()) if t.numel() > 1 else True

    def checksum(t):
        return int((t.to(torch.int64) & 0xFFFFFFFF).sum())

    if world > 1:
        # configs[4] as stated: 8 * 10^8 keys IN TOTAL, split evenly over the ranks
        total = 800_000_000
        n = total // world
        from vkradixsort_b200.dist import DistributedSorter

        keys = gen(n, 900 + rank)
        b0, b1 = torch.empty_like(keys), torch.empty_like(keys)
        ds = sorter if int(n * 1.25) + 1024 <= sorter.capacity else DistributedSorter(handle, n, world, rank, dev, pairs=False)
        ms, res = timed(lambda: ds.sort(b0, b1), lambda: b0.copy_(keys))
        sig = torch.tensor([res.numel(), checksum(res), n, checksum(keys)], dtype=torch.int64, device=dev)
        dist.all_reduce(sig)
        edges = torch.stack([(res[0] ^ FLIP).to(torch.int64), (res[-1] ^ FLIP).to(torch.int64)])
        gathered = [torch.empty_like(edges) for _ in range(world)]
        dist.all_gather(gathered, edges)
        ok = is_sorted(res) and int(sig[0]) == int(sig[2]) and int(sig[1]) == int(sig[3]) and all(bool(a[1] <= b[0]) for a, b in zip(gathered[:-1], gathered[1:]))
        out["strong_8e8"] = {"keys_total": total, "keys_per_gpu": n, "ms": ms, "mkeys_s": total / ms / 1e3, "verified": bool(ok),
                             "workload": "BASELINE.json configs[4]: 8*10^8 uniform uint32 keys in total, bucket exchange"}
        if ds is not sorter:
            ds.close()
        return out

    # ---- configs[2]: 10^8 key + payload pairs (payload = original index: stability is checkable) ----
    n = 100_000_000
    keys = gen(n, 11)
    vals = torch.arange(n, dtype=torch.int32, device=dev)
    k0, k1, v0, v1 = torch.empty_like(keys), torch.empty_like(keys), torch.empty_like(vals), torch.empty_like(vals)
    pc = capi.multi_push_constants(n, 32)

    def restore_pairs():
        k0.copy_(keys)
        v0.copy_(vals)

    ms, _ = timed(lambda: handle.multi_sort_pairs(k0, k1, v0, v1, None, pc), restore_pairs)
    ok = is_sorted(k0) and bool((keys[v0.long()] == k0).all())
    eq = k0[1:] == k0[:-1]  # stable: among equal keys the payloads (original indices) ascend
    ok = ok and bool((v0[1:][eq] > v0[:-1][eq]).all())
    out["pairs_1e8"] = {"n": n, "ms": ms, "mkeys_s": n / ms / 1e3, "verified": bool(ok), "algorithmic_bytes_per_pair": 96,
                        "gbs_96B_formula": n * 96 / (ms * 1e-3) / 1e9,
                        "workload": "BASELINE.json configs[2]: 10^8 uint32 key + uint32 payload pairs, stable"}
    del v0, v1, vals, eq

    # ---- the literal four-pass LSD loop of the north star, same keys as configs[1] ----
    def restore_keys():
        k0.copy_(keys)

    handle.set_schedule(capi.SCHEDULE_LSD)
    ms, _ = timed(lambda: handle.multi_sort(k0, k1, None, pc), restore_keys)
    handle.set_schedule(capi.SCHEDULE_AUTO)
    out["lsd_1e8"] = {"n": n, "ms": ms, "mkeys_s": n / ms / 1e3, "verified": is_sorted(k0) and checksum(k0) == checksum(keys),
                      "workload": "10^8 uniform uint32 keys, schedule LSD: four stable digit passes (segment histogram + segmented scatter)"}

    # ---- 64-bit keys: the reference's SORT_TYPE uint64_t variant, 44-bit test range (MultiRadixSort.cpp:128) ----
    n64 = 50_000_000
    g = torch.Generator(device=dev)
    g.manual_seed(12)
    k64 = torch.randint(0, 0x0FFFFFFFFFFF, (n64,), dtype=torch.int64, device=dev, generator=g)
    a64, b64 = torch.empty_like(k64), torch.empty_like(k64)
    pc64 = capi.multi_push_constants(n64, 32)
    ms, _ = timed(lambda: handle.multi_sort_u64(a64, b64, None, pc64), lambda: a64.copy_(k64))
    ok = bool((a64[1:] >= a64[:-1]).all()) and int(a64.sum()) == int(k64.sum())
    out["u64_5e7"] = {"n": n64, "ms": ms, "mkeys_s": n64 / ms / 1e3, "verified": bool(ok), "algorithmic_bytes_per_key": 256,
                      "gbs_256B_formula": n64 * 256 / (ms * 1e-3) / 1e9,
                      "workload": "5*10^7 uint64 keys below 2^44 (the reference's 64-bit variant: 8 digit passes)"}
    del k64, a64, b64

    # ---- configs[3]: N sweep 10^2 ... 10^8, single vs multi (28-bit keys, as the reference's published sweep) ----
    sweep = []
    for e in range(2, 9):
        m = 10 ** e
        src = keys[:m] & 0x0FFFFFFF
        a, b = k0[:m], k1[:m]
        pcm = capi.multi_push_constants(m, 32)
        row = {"n": m}
        reps = 5 if m <= 10 ** 7 else 3
        if m <= 10 ** 7:  # one work group sorting 10^8 keys takes most of a second: left out
            ms, _ = timed(lambda: handle.single_sort(a, b, capi.SinglePushConstants(m)), lambda: a.copy_(src), reps=reps)
            row["single_ms"] = ms
            row["single_ok"] = is_sorted(a)
        ms, _ = timed(lambda: handle.multi_sort(a, b, None, pcm), lambda: a.copy_(src), reps=reps)
        row["multi_ms"] = ms
        ms, _ = timed(lambda: handle.sort_auto(a, b, m), lambda: a.copy_(src), reps=reps)
        row["auto_ms"] = ms
        row["verified"] = is_sorted(a) and checksum(a) == checksum(src) and row.get("single_ok", True)
        sweep.append(row)
    out["nsweep"] = {"rows": sweep, "verified": all(r["verified"] for r in sweep),
                     "workload": "BASELINE.json configs[3]: N = 10^2 ... 10^8 reference-distribution (28-bit) keys; single = vkrs_single_sort "
                                 "(one work group), multi = vkrs_multi_sort, auto = vkrs_sort_auto (the dispatcher)"}
    del k0, k1, keys

    # ---- configs[4] on one GPU: 8 * 10^8 keys ----
    n8 = 800_000_000
    keys8 = gen(n8, 13)
    a8, b8 = torch.empty_like(keys8), torch.empty_like(keys8)
    pc8 = capi.multi_push_constants(n8, 32)
    ms, _ = timed(lambda: handle.multi_sort(a8, b8, None, pc8), lambda: a8.copy_(keys8), reps=2)
    out["keys_8e8"] = {"n": n8, "ms": ms, "mkeys_s": n8 / ms / 1e3, "verified": is_sorted(a8) and checksum(a8) == checksum(keys8),
                       "schedule_resolved": capi.schedule_name(handle.resolve_schedule(n8)), "bucket_schedule": handle.bucket_stats(),
                       "workload": "BASELINE.json configs[4] at one GPU: 8*10^8 uniform uint32 keys"}
    return out


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_b200(args) -> int:
    import torch

    from vkradixsort_b200 import Handle, capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    n = args.n
    pairs = args.workload == "pairs"
    steps, warmup = args.steps, max(3, args.warmup)
    stream = torch.cuda.current_stream()
    handle = Handle(local_rank, n)
    if args.variant is not None:
        handle.set_variant(args.variant)
    if args.schedule is not None:
        handle.set_schedule(args.schedule)

    host_keys = make_keys(n, SEED + rank)
    pristine = torch.from_numpy(host_keys.view(np.int32)).to(dev)
    buf0 = torch.empty_like(pristine)
    buf1 = torch.empty_like(pristine)
    if pairs:
        val_pristine = torch.arange(n, dtype=torch.int32, device=dev)
        val0, val1 = torch.empty_like(val_pristine), torch.empty_like(val_pristine)
    pc = capi.multi_push_constants(n, 32)

    if world > 1:
        from vkradixsort_b200.dist import DistributedSorter

        sorter = DistributedSorter(handle, n, world, rank, dev, pairs=False)

    def restore():
        buf0.copy_(pristine)  # untimed; also evicts the previous step's output from L2 (400 MB > 126 MB)
        if pairs:
            val0.copy_(val_pristine)

    def one_sort():
        if world > 1:
            return sorter.sort(buf0, buf1)
        if pairs:
            handle.multi_sort_pairs(buf0, buf1, val0, val1, None, pc)
        else:
            handle.multi_sort(buf0, buf1, None, pc)
        return buf0

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: `value` ----
    for _ in range(warmup):
        restore()
        out = one_sort()
    handle.check_device_error()
    barrier()
    launches0 = handle.launch_count
    sampler = ClockSampler(local_rank).start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        restore()
        if dist is not None:
            dist.barrier()
        starts[i].record(stream)
        out = one_sort()
        ends[i].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = handle.launch_count - launches0
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / steps
    handle.check_device_error()
    main_bucket_stats = handle.bucket_stats() if not pairs else None  # control words of the last timed sort (later configs overwrite them)

    # ---- verification of the last timed output (size-independent properties, on the device) ----
    # sortedness per rank + boundary order between ranks + the multiset: count, sum, XOR and the 256 top-byte bucket
    # counts of output and input must agree (a swapped pair of keys with equal sum changes the XOR or a bucket count).
    flip = torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)
    o = out ^ flip
    sorted_ok = bool((o[1:] >= o[:-1]).all()) if o.numel() > 1 else True

    def multiset_signature(t):
        t64 = t.to(torch.int64) & 0xFFFFFFFF
        x = t.clone()
        while x.numel() > 1:  # XOR reduction on the device (torch has no bitwise reduce)
            h = x.numel() // 2
            rest = x[2 * h:]
            x = torch.cat([x[:h] ^ x[h:2 * h], rest])
        xor = (x.to(torch.int64) & 0xFFFFFFFF) if x.numel() else torch.zeros(1, dtype=torch.int64, device=dev)
        buckets = torch.bincount((t64 >> 24).to(torch.int64), minlength=256)
        return torch.cat([torch.tensor([t.numel()], dtype=torch.int64, device=dev), t64.sum().reshape(1), buckets]), xor.reshape(1)

    sig_out, xor_out = multiset_signature(out)
    sig_in, xor_in = multiset_signature(pristine)
    if dist is not None:
        for t in (sig_out, sig_in):
            dist.all_reduce(t)
        xs = [torch.empty_like(xor_out) for _ in range(world)]
        dist.all_gather(xs, xor_out)
        xor_out = xs[0].clone()
        for x in xs[1:]:
            xor_out ^= x
        xs = [torch.empty_like(xor_in) for _ in range(world)]
        dist.all_gather(xs, xor_in)
        xor_in = xs[0].clone()
        for x in xs[1:]:
            xor_in ^= x
        # boundary order between consecutive ranks
        edges = torch.stack([o[0].to(torch.int64), o[-1].to(torch.int64)]) if o.numel() else torch.zeros(2, dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(edges) for _ in range(world)]
        dist.all_gather(gathered, edges)
        for a, b in zip(gathered[:-1], gathered[1:]):
            sorted_ok = sorted_ok and bool(a[1] <= b[0])
    verified = sorted_ok and bool(torch.equal(sig_out, sig_in)) and bool(torch.equal(xor_out, xor_in)) and int(sig_out[0]) == n * world
    if pairs and world == 1:
        # stable: among equal keys payloads (original indices) ascend; keys[payload] reproduces the output
        verified = verified and bool((pristine[val0.long()] == buf0).all())
    del o

    # ---- per-kernel timing of the dominant kernel: `roofline` (this rank's kernels; printed by rank 0) ----
    peak, peak_src = load_peaks()
    roofline = None
    if True:
        handle.set_profiling(True)
        for _ in range(min(steps, 20)):
            restore()
            one_sort()
        torch.cuda.synchronize()
        prof = handle.profile()
        handle.set_profiling(False)
        reps = max(1, min(steps, 20))
        big = {k: v for k, v in prof.items() if algo_bytes_per_key(k, pairs) > 0 and v["ms"] / v["launches"] > 0.02}  # gated-off passes take ~5 us
        dom = max(big.items(), key=lambda kv: kv[1]["ms"])
        total_prof = sum(v["ms"] for v in prof.values())
        bytes_per_key = algo_bytes_per_key(dom[0], pairs)
        dom_ms = dom[1]["ms"] / dom[1]["launches"]
        achieved = n * bytes_per_key / (dom_ms * 1e-3) / 1e9
        sort_bytes = (96 if pairs else ALGO_BYTES_PER_KEY_SORT)
        kernels = {}
        for k, v in big.items():
            per_launch = v["ms"] / v["launches"]
            gbs = n * algo_bytes_per_key(k, pairs) / (per_launch * 1e-3) / 1e9
            kernels[k] = {"launches_per_sort": v["launches"] / reps, "avg_launch_ms": per_launch, "achieved": gbs, "frac": gbs / peak,
                          "share_of_step": v["ms"] / total_prof if total_prof else None}
        roofline = {
            "bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": n * bytes_per_key, "avg_launch_ms": dom_ms,
            "share_of_step": dom[1]["ms"] / total_prof if total_prof else None,
            "kernels": kernels,
            "kernels_ms_per_sort": {k: v["ms"] / reps for k, v in prof.items()},
            "kernels_note": "per-kernel times come from a separate profiling loop (one event pair per launch): that defeats the "
                            "programmatic-dependent-launch overlap between consecutive kernels, so they add up to a little more than ms_per_step",
            "whole_sort": {"formula_bytes_per_key": sort_bytes,
                           "achieved_gbs": n * sort_bytes / (ms_per_step * 1e-3) / 1e9,
                           "frac_of_measured_peak": n * sort_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                           "frac_of_nominal_8TBs": n * sort_bytes / (ms_per_step * 1e-3) / 1e9 / 8000.0},
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                roofline["traffic"] = json.load(open(traffic_file)).get(dom[0].split("<")[0])
                roofline["traffic_source"] = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` "
                                              "capture of this kernel at this workload (a committed number, not measured in this run)")
            except Exception:
                pass

    # ---- end to end with HOST buffers: `e2e` ----
    # N = 1: the host-buffer C-ABI call vkrs_multi_sort_host (pinned host keys -> H2D -> sort -> D2H), wall clock around the
    #        call, split into its three stages by the library's own events (vkrs_host_timings).
    # N > 1: every rank copies its pinned host keys to the device, the global sort runs, the rank's sorted range comes
    #        back to pinned host memory; wall clock between barriers, max over ranks.
    e2e = None
    if not pairs:
        pinned_src = torch.from_numpy(host_keys.view(np.int32)).pin_memory()
        e2e_steps = min(steps, 10)
        t_total, split = 0.0, {"h2d_ms": 0.0, "sort_ms": 0.0, "d2h_ms": 0.0}
        if world == 1:
            pinned = torch.empty_like(pinned_src).pin_memory()
            for i in range(2 + e2e_steps):
                pinned.copy_(pinned_src)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                handle.multi_sort_host(pinned, n)  # returns with the sorted keys back in host memory
                t1 = time.perf_counter()
                if i >= 2:
                    t_total += t1 - t0
                    for k, v in handle.host_timings().items():
                        split[k] += v / e2e_steps
            res = pinned.numpy().view(np.uint32)
            e2e_ok = bool(np.all(res[1:] >= res[:-1])) and int(res.astype(np.uint64).sum()) == int(host_keys.astype(np.uint64).sum())
            d2h_bytes = 4 * n
            api = "vkrs_multi_sort_host (pinned host buffer in, sorted in place)"
        else:
            pinned_out = torch.empty(sorter.capacity, dtype=torch.int32).pin_memory()
            got = 0
            for i in range(2 + e2e_steps):
                barrier()
                t0 = time.perf_counter()
                buf0.copy_(pinned_src, non_blocking=True)
                res_dev = sorter.sort(buf0, buf1)
                got = int(res_dev.numel())
                pinned_out[:got].copy_(res_dev, non_blocking=True)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                if i >= 2:
                    t_total += t1 - t0
            t = torch.tensor([t_total], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_total = float(t.item())
            res = pinned_out[:got].numpy().view(np.uint32)
            e2e_ok = bool(np.all(res[1:] >= res[:-1])) if got > 1 else True
            d2h_bytes = 4 * n  # on average: every rank receives n keys of the N * n
            api = "pinned host keys -> device, DistributedSorter.sort, sorted range -> pinned host (per rank; max over ranks)"
            split = None
        verified = verified and e2e_ok
        e2e_ms = 1e3 * t_total / e2e_steps
        e2e = {"value": n * world / (e2e_ms * 1e-3) / 1e6, "unit": "Mkeys/s", "h2d_bytes_per_step": 4 * n,
               "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms, "steps": e2e_steps, "api": api}
        if split:
            e2e["stages_ms"] = {**split, "host_overhead_ms": e2e_ms - sum(split.values())}
            e2e["pcie_gbs"] = {"h2d": 4 * n / (split["h2d_ms"] * 1e-3) / 1e9 if split["h2d_ms"] else None,
                               "d2h": 4 * n / (split["d2h_ms"] * 1e-3) / 1e9 if split["d2h_ms"] else None}

    # ---- CPU baseline: the reference's own CPU arm, same keys (rank 0, N=1) ----
    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as O  # cpu_baseline leg only

        O.build()
        k = host_keys.copy()
        std_ms = O.std_sort(k)
        cpu_ok = bool(np.array_equal(k[:: 1009], np.sort(host_keys)[:: 1009])) if n <= 2 * 10**7 else True
        cpu_baseline = {"value": n / std_ms / 1e3, "unit": "Mkeys/s", "cores": 1, "kind": "port",
                        "sample": f"all {n} keys of the step, in-place std::sort as MultiRadixSort::sort "
                                  f"(MultiRadixSort.cpp:141-146), {std_ms:.0f} ms",
                        "host_cpus": os.cpu_count(), "ok": cpu_ok}
        # the GPU output of the same keys must equal the CPU-sorted keys element-wise (testSort)
        restore()
        res = one_sort().cpu().numpy().view(np.uint32)
        verified = verified and O.test_sort(k, res) == -1

    # ---- the other BASELINE.json configurations, measured after the headline so that it is unchanged: `configs` ----
    extra = None
    if not pairs and not args.no_configs:
        extra = other_configs(torch, handle, capi, dev, rank, world, dist, sorter if world > 1 else None)
        verified = verified and all(v.get("verified", True) for v in extra.values() if isinstance(v, dict))

    if rank == 0:
        total_keys = n * world
        resolved = handle.resolve_schedule(n) if not pairs else capi.SCHEDULE_LSD
        line = {
            "metric": METRIC, "value": total_keys / (ms_per_step * 1e-3) / 1e6, "unit": "Mkeys/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {**common_config(pairs, world, n), "l2": "inputs (400 MB) larger than L2 (126 MB); input restored by a "
                       "400 MB device copy before every step",
                       "schedule": capi.schedule_name(handle.schedule),
                       "schedule_resolved": capi.schedule_name(resolved),
                       "kernels": ("msd_piece_histogram_kernel + msd_scatter_kernel (x2, unstable top-digit passes) + msd_local_tile_kernel"
                                   if resolved == capi.SCHEDULE_BUCKET else
                                   "segment_histogram_kernel + segmented_scatter_kernel per digit, tile variant " + capi.variant_name(handle.variant)),
                       "lsd_tile_variant": capi.variant_name(handle.variant) if resolved != capi.SCHEDULE_BUCKET else "not run (bucket schedule)",
                       "timing": "CUDA events on the launch stream around each sort, summed; max over ranks"},
            "clocks": clocks, "gpu_launches": int(launches), "verified": bool(verified),
        }
        if not pairs:
            line["config"]["bucket_schedule"] = main_bucket_stats  # shifts, fallback flag, largest bucket, big buckets of the last timed sort
        if roofline:
            line["roofline"] = roofline
        if e2e:
            line["e2e"] = e2e
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if extra:
            line["configs"] = extra
            if "lsd_1e8" in extra:
                line["lsd_ms"] = extra["lsd_1e8"]["ms"]  # the north star's literal four-pass LSD loop, beside the headline
        print(json.dumps(line), file=json_out, flush=True)
    handle.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0 if verified else 1


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="keys", choices=["keys", "pairs"])
    ap.add_argument("--n", type=int, default=N_KEYS, help="keys per GPU (default: the BASELINE configuration)")
    ap.add_argument("--variant", type=int, default=None, help="tuning: kernel tile variant of the stable LSD passes (forces the LSD schedule)")
    ap.add_argument("--schedule", type=int, default=None, help="tuning: 1 = LSD, 2 = LSD with unstable first pass, 3 = bucket (default: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (pairs, u64, LSD, N sweep, 8*10^8 keys)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
