"""compute-sanitizer target: one bucket-schedule sort of a bucket_probe distribution.  python tools/san_case.py <dist> <n> [schedule]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv, args = sys.argv[:1], sys.argv[1:]
os.environ.setdefault("PROBE_QUICK", "1")
import importlib.util
spec = importlib.util.spec_from_file_location("bp", os.path.join(os.path.dirname(os.path.abspath(__file__)), "bucket_probe.py"))
bp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bp)
from vkradixsort_b200 import Handle, capi
dist, n = args[0], int(float(args[1]))
sched = int(args[2]) if len(args) > 2 else capi.SCHEDULE_BUCKET
h = Handle(0, n)
h.set_schedule(sched)
keys = bp.gen(dist, n, 1000 + n)
want = bp.expect_sorted(keys)
b0, b1 = keys.clone(), torch.full_like(keys, 0x5A5A5A5A)
h.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32))
torch.cuda.synchronize()
print("SAN_CASE", dist, n, "bad:", bp.first_bad(b0, want), h.bucket_stats() if sched == capi.SCHEDULE_BUCKET else None)
