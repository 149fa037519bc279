"""BASELINE.json config 4: N sweep 10^2 ... 10^8 uint32 keys, single vs multi path, one B200.
Device time per sort (CUDA events around the call, median of reps; input restored before each rep).
    python tools/nsweep.py > profiles/r01_nsweep.jsonl
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi  # noqa: E402

dev = torch.device("cuda:0")
h = Handle(0, 10**8)
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


for e in range(2, 9):
    n = 10**e
    keys = np.random.default_rng(e).integers(0, 1 << 28, size=n, dtype=np.uint32)  # the reference's 28-bit range
    pristine = torch.from_numpy(keys.view(np.int32)).to(dev)
    b0, b1 = torch.empty_like(pristine), torch.empty_like(pristine)
    expect = np.sort(keys)
    reps = 30 if n <= 10**6 else 10
    row = {"n": n}
    pc = capi.multi_push_constants(n, 32)
    row["multi_ms"] = timed(lambda: h.multi_sort(b0, b1, None, pc), lambda: b0.copy_(pristine), reps)
    assert np.array_equal(b0.cpu().numpy().view(np.uint32), expect)
    if n <= 10**6:
        spc = capi.SinglePushConstants(n)
        row["single_ms"] = timed(lambda: h.single_sort(b0, b1, spc), lambda: b0.copy_(pristine), reps)
        assert np.array_equal(b0.cpu().numpy().view(np.uint32), expect)
    row["auto_ms"] = timed(lambda: h.sort_auto(b0, b1, n), lambda: b0.copy_(pristine), reps)
    row["staged_nb32_ms"] = timed(lambda: h.multi_sort_staged(b0, b1, torch.empty(max(1, pc.g_num_workgroups) * 256, dtype=torch.int32, device=dev), pc),
                                  lambda: b0.copy_(pristine), 5) if n <= 10**7 else None
    row["multi_mkeys_s"] = n / row["multi_ms"] / 1e3
    print(json.dumps(row), flush=True)
