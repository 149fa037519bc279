"""Crossover of the keys-only schedules (vkrs_set_schedule) over N, one B200: device time per sort
(CUDA events, median of reps, input restored before each rep), uniform 32-bit and the reference's 28-bit keys.
    python tools/schedule_sweep.py > profiles/r01_schedule_sweep.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi  # noqa: E402

dev = torch.device("cuda:0")
h = Handle(0, 1 << 28)
stream = torch.cuda.current_stream()


def timed(fn, restore, reps):
    ts = []
    for i in range(reps + 3):
        restore()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


sizes = [1 << 20, 1 << 22, 1 << 24, 1 << 25, 40_000_000, 50_000_000, 1 << 26, 80_000_000, 100_000_000, 1 << 27, 200_000_000, 1 << 28]
for mask_bits in (32, 28):
    for n in sizes:
        g = torch.Generator(device=dev)
        g.manual_seed(n)
        pristine = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
        if mask_bits < 32:
            pristine &= (1 << mask_bits) - 1
        b0, b1 = torch.empty_like(pristine), torch.empty_like(pristine)
        pc = capi.multi_push_constants(n, 32)
        row = {"n": n, "key_bits": mask_bits}
        for sched, name in ((capi.SCHEDULE_LSD, "lsd"), (capi.SCHEDULE_LSD_UNSTABLE_FIRST, "lsd_unstable_first"), (capi.SCHEDULE_BUCKET, "bucket"),
                            (capi.SCHEDULE_AUTO, "auto")):
            h.set_schedule(sched)
            row[name + "_ms"] = round(timed(lambda: h.multi_sort(b0, b1, None, pc), lambda: b0.copy_(pristine), 10), 4)
            x = b0 ^ -(1 << 31)
            assert bool((x[1:] >= x[:-1]).all()), (n, name)
            if sched == capi.SCHEDULE_BUCKET:
                row["bucket_stats"] = h.bucket_stats()
        row["best"] = min(("lsd", "lsd_unstable_first", "bucket"), key=lambda k: row[k + "_ms"])
        row["auto_gkeys_s"] = round(n / row["auto_ms"] / 1e6, 2)
        print(json.dumps(row), flush=True)
        del pristine, b0, b1
