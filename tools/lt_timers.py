"""Phase timers of msd_local_tile_kernel (library built with -DVKRS_LT_TIMERS): cycles of thread 0 per phase, summed over CTAs.
    VKRS_LIB_PATH=... python tools/lt_timers.py [n]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(5)
keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
b0, b1 = torch.empty_like(keys), torch.empty_like(keys)
h = Handle(0, n)
h.set_schedule(capi.SCHEDULE_BUCKET)
pc = capi.multi_push_constants(n, 32)
for _ in range(2):
    b0.copy_(keys); h.multi_sort(b0, b1, None, pc)
torch.cuda.synchronize()
h.debug_counters(True)
b0.copy_(keys); h.multi_sort(b0, b1, None, pc)
torch.cuda.synchronize()
c = h.debug_counters(True)
names = {0: "top barrier wait", 1: "prefetch issue + store prev item", 2: "zero + barrier", 3: "count", 4: "barrier after count",
         5: "scan (3 barriers)", 6: "place", 7: "barrier after place", 8: "fix-up", 9: "cp.async wait", 11: "loop tail (robust path etc.)"}
ctas = max(1, c[10])
tot = sum(c[i] for i in names)
print(json.dumps({"ctas": ctas, "cycles_per_cta": tot // ctas, "phases": {names[i]: [c[i] // ctas, round(100.0 * c[i] / max(1, tot), 1)] for i in sorted(names)}}, indent=1))
