"""GPU-box probe of the keys-only schedules (vkrs_set_schedule): a correctness matrix with per-stage
diagnosis, then per-kernel device times at full size.
    python tools/bucket_probe.py [n_big] [reps]
Writes one JSON object per line to stdout (the caller redirects it into gpurun_out/).
torch.sort is only the checker here, never part of the product path.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi  # noqa: E402

n_big = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
FLIP = -(1 << 31)


def out(**kw):
    print(json.dumps(kw), flush=True)


def expect_sorted(t):
    """ascending in unsigned order, computed on the device"""
    return torch.sort(t ^ FLIP).values ^ FLIP


def gen(name, n, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    r = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    if name == "uniform32":
        return r
    if name == "reference28":
        return r & 0x0FFFFFFF
    if name == "bits20":
        return r & 0x000FFFFF
    if name == "bits12":
        return r & 0xFFF
    if name == "sorted":
        return expect_sorted(r)
    if name == "all_equal":
        return torch.full_like(r, 0x1EADBEEF)
    if name == "hot_prefix":
        return torch.where(r > 0, (r & 0xFFFF) | 0x2BCD0000, r)
    if name == "dup1024":
        return (r & 1023) * 4194301
    raise KeyError(name)


def first_bad(a, b):
    bad = (a != b).nonzero()
    if bad.numel() == 0:
        return None
    i = int(bad[0])
    return {"index": i, "mismatches": int(bad.numel()), "got": int(a[i]) & 0xFFFFFFFF, "want": int(b[i]) & 0xFFFFFFFF}


def diagnose(h, keys, want):
    """Re-run the bucket schedule stage by stage and say which stage breaks first."""
    n = keys.numel()
    pc = capi.multi_push_constants(n, 32)
    rep = {}
    for stage in (1, 2, 3):
        b0, b1 = keys.clone(), torch.full_like(keys, 0x5A5A5A5A)
        h.debug_bucket_stop(stage)
        h.multi_sort(b0, b1, None, pc)
        torch.cuda.synchronize()
        st = h.bucket_stats()
        res = b1 if stage == 1 else b0
        base = st["key_min"] if st["recount"] else 0  # the digits are taken from key - base
        perm_ok = bool(torch.equal(expect_sorted(res), want))
        if stage == 1:
            d = ((res - base) >> st["shift1"]) & 255
            grouped = bool((d[1:] >= d[:-1]).all())
        elif stage == 2:
            d = (((res - base) >> st["shift2"]) & 0xFFFF).to(torch.int64)
            grouped = bool((d[1:] >= d[:-1]).all())
        else:
            grouped = first_bad(res, want) is None
        rep[f"stage{stage}"] = {"permutation": perm_ok, "grouped_or_sorted": grouped, "stats": st}
    h.debug_bucket_stop(0)
    return rep


def correctness(h):
    ok_all = True
    sizes = [1, 33, 6143, 6145, 100_003, 1_000_001, 20_000_003]
    names = ["uniform32", "reference28", "bits20", "bits12", "sorted", "all_equal", "hot_prefix", "dup1024"]
    for sched in (capi.SCHEDULE_LSD_UNSTABLE_FIRST, capi.SCHEDULE_BUCKET):
        h.set_schedule(sched)
        for n in sizes:
            for name in names:
                keys = gen(name, n, 1000 + n)
                want = expect_sorted(keys)
                b0, b1 = keys.clone(), torch.full_like(keys, 0x5A5A5A5A)
                try:
                    h.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32))
                    torch.cuda.synchronize()
                    h.check_device_error()
                    bad = first_bad(b0, want)
                except Exception as e:  # a CUDA error poisons the context: report and stop
                    out(kind="error", schedule=sched, n=n, dist=name, error=repr(e))
                    return False
                if bad is not None:
                    ok_all = False
                    row = dict(kind="mismatch", schedule=sched, n=n, dist=name, bad=bad)
                    if sched == capi.SCHEDULE_BUCKET:
                        row["stats"] = h.bucket_stats()
                        row["stages"] = diagnose(h, keys, want)
                        h.set_schedule(sched)
                    out(**row)
        out(kind="correctness", schedule=sched, name=capi.schedule_name(sched), ok=ok_all)
    return ok_all


def timing(h, name, n, schedules):
    keys = gen(name, n, 7)
    want = expect_sorted(keys)
    b0, b1 = torch.empty_like(keys), torch.empty_like(keys)
    pc = capi.multi_push_constants(n, 32)
    for sched in schedules:
        h.set_schedule(sched)
        for _ in range(3):
            b0.copy_(keys)
            h.multi_sort(b0, b1, None, pc)
        torch.cuda.synchronize()
        ok = first_bad(b0, want) is None
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in ev:
            b0.copy_(keys)
            a.record()
            h.multi_sort(b0, b1, None, pc)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        h.set_profiling(True)
        for _ in range(3):
            b0.copy_(keys)
            h.multi_sort(b0, b1, None, pc)
        prof = h.profile()
        h.set_profiling(False)
        out(kind="timing", dist=name, n=n, schedule=sched, name=capi.schedule_name(sched), ok=ok,
            ms_median=round(ms[len(ms) // 2], 4), ms_best=round(ms[0], 4), gkeys_s=round(n / ms[len(ms) // 2] / 1e6, 2),
            stats=h.bucket_stats() if sched == capi.SCHEDULE_BUCKET else None,
            kernels_us={k: round(1e3 * v["ms"] / v["launches"], 2) for k, v in prof.items()},
            launches_per_sort={k: v["launches"] // 3 for k, v in prof.items()})


def main():
    t0 = time.time()
    h = Handle(0, n_big)
    if os.environ.get("PROBE_QUICK"):
        if os.environ.get("PROBE_SKIP_CORRECTNESS") or correctness(h):
            timing(h, "uniform32", n_big, [capi.SCHEDULE_BUCKET])
            if not os.environ.get("PROBE_ONLY_UNIFORM"):
                timing(h, "reference28", n_big, [capi.SCHEDULE_BUCKET])
                timing(h, "hot_prefix", n_big, [capi.SCHEDULE_BUCKET])
                timing(h, "sorted", n_big, [capi.SCHEDULE_BUCKET])
                timing(h, "dup1024", n_big, [capi.SCHEDULE_BUCKET])
        out(kind="done", seconds=round(time.time() - t0, 1))
        return
    if correctness(h):
        timing(h, "uniform32", n_big, [capi.SCHEDULE_LSD, capi.SCHEDULE_LSD_UNSTABLE_FIRST, capi.SCHEDULE_BUCKET])
        timing(h, "reference28", n_big, [capi.SCHEDULE_BUCKET])
        timing(h, "sorted", n_big, [capi.SCHEDULE_LSD_UNSTABLE_FIRST, capi.SCHEDULE_BUCKET])
        timing(h, "dup1024", n_big, [capi.SCHEDULE_BUCKET])
        for n in (1 << 22, 1 << 24, 1 << 25, 1 << 26):
            timing(h, "uniform32", n, [capi.SCHEDULE_LSD, capi.SCHEDULE_LSD_UNSTABLE_FIRST, capi.SCHEDULE_BUCKET])
    out(kind="done", seconds=round(time.time() - t0, 1))


if __name__ == "__main__":
    main()
