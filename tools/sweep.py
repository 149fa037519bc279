"""Tuning aid (GPU box): per-kernel device time of the fused sort for every precompiled tile variant.
    python tools/sweep.py [n] [reps] [variant,variant,...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
only = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
dev = torch.device("cuda:0")
keys = np.random.default_rng(1).integers(0, 1 << 32, size=n, dtype=np.uint32)
pristine = torch.from_numpy(keys.view(np.int32)).to(dev)
b0, b1 = torch.empty_like(pristine), torch.empty_like(pristine)
pc = capi.multi_push_constants(n, 32)
h = Handle(0, n)
rows = []
for v in (only if only is not None else range(capi.num_variants())):
    h.set_variant(v)
    for _ in range(2):
        b0.copy_(pristine)
        h.multi_sort(b0, b1, None, pc)
    h.set_profiling(True)
    h.debug_counters(True)
    for _ in range(reps):
        b0.copy_(pristine)
        h.multi_sort(b0, b1, None, pc)
    prof = h.profile()
    dbg = h.debug_counters(False)
    h.set_profiling(False)
    h.check_device_error()
    ok = bool((b0[1:] ^ -(1 << 31) >= b0[:-1] ^ -(1 << 31)).all())
    row = {"variant": v, "name": capi.variant_name(v), "ok": ok,
           **{k: round(x["ms"] / x["launches"], 4) for k, x in prof.items()}}
    row["sort_ms"] = round(sum(x["ms"] for x in prof.values()) / reps, 4)
    if dbg[4]:
        t = dbg[4]  # tiles processed by all control warps
        row["ctrl_cyc_per_tile"] = {"wait_counts": dbg[0] // t, "claim": dbg[1] // t, "lookback": dbg[2] // t,
                                    "polls": round(dbg[3] / t, 2)}
        names = ["wait_tile", "rank", "barA", "digits", "wait_prefix", "write_out", "barB", "scatter"]
        row["worker_cyc_per_tile"] = {n: dbg[8 + i] // t for i, n in enumerate(names)}
    rows.append(row)
    print(json.dumps(row), flush=True)
