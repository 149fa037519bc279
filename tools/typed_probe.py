"""Device time of vkrs_multi_sort_typed (uint32 / int32 / float32 keys) per schedule, one B200.
    python tools/typed_probe.py [n] [reps]
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
bits = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
floats = (torch.randn(n, device=dev, generator=g) * 1e6).view(torch.int32)
h = Handle(0, n)
pc = capi.multi_push_constants(n, 32)
b0, b1 = torch.empty_like(bits), torch.empty_like(bits)
for name, src, kt in (("uint32", bits, capi.KEY_U32), ("int32", bits, capi.KEY_I32), ("float32", floats, capi.KEY_F32)):
    for sched in (capi.SCHEDULE_LSD, capi.SCHEDULE_AUTO, capi.SCHEDULE_BUCKET):
        h.set_schedule(sched)
        ts = []
        for i in range(reps + 3):
            b0.copy_(src)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); h.multi_sort_typed(b0, b1, None, pc, kt); b.record(); torch.cuda.synchronize()
            if i >= 3: ts.append(a.elapsed_time(b))
        if kt == capi.KEY_F32:
            ok = bool((b0.view(torch.float32)[1:] >= b0.view(torch.float32)[:-1]).all())
        elif kt == capi.KEY_I32:
            ok = bool((b0[1:] >= b0[:-1]).all())
        else:
            x = b0 ^ -(1 << 31); ok = bool((x[1:] >= x[:-1]).all())
        ts.sort()
        print(json.dumps({"keys": name, "n": n, "schedule": capi.schedule_name(sched), "ms_median": round(ts[len(ts) // 2], 4), "sorted": ok,
                          "stats": h.bucket_stats() if sched == capi.SCHEDULE_BUCKET else None}), flush=True)
