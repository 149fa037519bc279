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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O  # the one other place bench.py may execute oracle/

    O.build()
    threads = O.num_threads()
    sample = 10_000_000  # bounded sample of the 10^8-key workload per step
    nb = 512             # the reference's published best nb at 10^7 (README timings)
    pristine = make_keys(sample, SEED)
    for _ in range(args.warmup):
        O.multi_sort(pristine, nb)
    t_total = 0.0
    for _ in range(args.steps):
        keys = pristine.copy()
        t0 = time.perf_counter()
        out = O.multi_sort(keys, nb)[0]
        t_total += time.perf_counter() - t0
    assert np.all(out[1:] >= out[:-1])
    ms = 1e3 * t_total / args.steps
    mkeys = sample / (ms * 1e-3) / 1e6
    k1 = pristine.copy()
    std_ms = O.std_sort(k1)
    k2 = pristine.copy()
    par_ms = O.parallel_sort(k2)
    line = {
        "impl": "reference", "metric": METRIC, "value": mkeys, "unit": "Mkeys/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {**common_config(False, 1, N_KEYS), "sample_keys_per_step": sample},
        "cpu_baseline": {"value": mkeys, "unit": "Mkeys/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} of the 10^8 keys per step, restated multi_radixsort shaders "
                                   f"(oracle/vkrs_oracle.c, nb={nb}), OpenMP over work groups",
                         "also_mkeys_per_s": {"std_sort_1_thread": sample / std_ms / 1e3,
                                              f"gnu_parallel_sort_{threads}_threads": sample / par_ms / 1e3}},
        "e2e": {"value": mkeys, "unit": "Mkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the Vulkan reference cannot be built in this image (no Vulkan headers/loader, no glslc); "
                "this is its algorithm restated in C on the host cores",
    }
    print(json.dumps(line), flush=True)
    return 0


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

    # ---- verification of the last timed output (size-independent properties, on the device) ----
    flip = torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)
    o = out ^ flip
    sorted_ok = bool((o[1:] >= o[:-1]).all()) if o.numel() > 1 else True
    sum_out = out.to(torch.int64).sum()
    sum_in = pristine.to(torch.int64).sum()
    cnt = torch.tensor([out.numel()], dtype=torch.int64, device=dev)
    if dist is not None:
        for t in (sum_out, sum_in, cnt):
            dist.all_reduce(t)
        # boundary order between consecutive ranks
        edges = torch.stack([o[0].to(torch.int64), o[-1].to(torch.int64)]) if o.numel() else torch.zeros(2, dtype=torch.int64, device=dev)
        gathered = [torch.empty_like(edges) for _ in range(world)]
        dist.all_gather(gathered, edges)
        for a, b in zip(gathered[:-1], gathered[1:]):
            sorted_ok = sorted_ok and bool(a[1] <= b[0])
    verified = sorted_ok and int(sum_out) == int(sum_in) and int(cnt) == n * world
    if pairs and world == 1:
        # stable: among equal keys payloads (original indices) ascend; keys[payload] reproduces the output
        verified = verified and bool((pristine[val0.long()] == buf0).all())
    del o

    # ---- per-kernel timing of the dominant kernel: `roofline` (N=1 path only) ----
    peak, peak_src = load_peaks()
    roofline = None
    if world == 1:
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
            "whole_sort": {"formula_bytes_per_key": sort_bytes,
                           "achieved_gbs": n * sort_bytes / (ms_per_step * 1e-3) / 1e9,
                           "frac_of_measured_peak": n * sort_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                           "frac_of_nominal_8TBs": n * sort_bytes / (ms_per_step * 1e-3) / 1e9 / 8000.0},
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                roofline["traffic"] = json.load(open(traffic_file)).get(dom[0].split("<")[0])
            except Exception:
                pass

    # ---- end to end through the host-buffer entry point: `e2e` ----
    e2e = None
    if world == 1 and not pairs:
        pinned_src = torch.from_numpy(host_keys.view(np.int32)).pin_memory()
        pinned = torch.empty_like(pinned_src).pin_memory()
        e2e_steps = min(steps, 10)
        t_total = 0.0
        for i in range(2 + e2e_steps):
            pinned.copy_(pinned_src)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            handle.multi_sort_host(pinned, n)  # returns with the sorted keys back in host memory
            t1 = time.perf_counter()
            if i >= 2:
                t_total += t1 - t0
        res = pinned.numpy().view(np.uint32)
        e2e_ok = bool(np.all(res[1:] >= res[:-1])) and int(res.astype(np.uint64).sum()) == int(host_keys.astype(np.uint64).sum())
        verified = verified and e2e_ok
        e2e_ms = 1e3 * t_total / e2e_steps
        e2e = {"value": n / (e2e_ms * 1e-3) / 1e6, "unit": "Mkeys/s", "h2d_bytes_per_step": 4 * n,
               "d2h_bytes_per_step": 4 * n, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "api": "vkrs_multi_sort_host (pinned host buffer in, sorted in place)"}

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

    if rank == 0:
        total_keys = n * world
        line = {
            "metric": METRIC, "value": total_keys / (ms_per_step * 1e-3) / 1e6, "unit": "Mkeys/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {**common_config(pairs, world, n), "l2": "inputs (400 MB) larger than L2 (126 MB); input restored by a "
                       "400 MB device copy before every step", "variant": capi.variant_name(handle.variant),
                       "schedule": capi.schedule_name(handle.schedule),
                       "timing": "CUDA events on the launch stream around each sort, summed; max over ranks"},
            "clocks": clocks, "gpu_launches": int(launches), "verified": bool(verified),
        }
        if not pairs:
            line["config"]["bucket_schedule"] = handle.bucket_stats()  # shifts, fallback flag, largest bucket of the last bucket-schedule sort
        if roofline:
            line["roofline"] = roofline
        if e2e:
            line["e2e"] = e2e
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
