"""BASELINE.json config 5 at G=1: 8 x 10^8 keys on one GPU; sortedness + checksum on the device."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
dev = torch.device("cuda:0")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 800_000_000
g = torch.Generator(device=dev); g.manual_seed(5)
b0 = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
s_in = int(b0.to(torch.int64).sum())
b1 = torch.empty_like(b0)
h = Handle(0, n)
pc = capi.multi_push_constants(n, 32)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); h.multi_sort(b0, b1, None, pc); b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
flip = torch.tensor(-(1 << 31), dtype=torch.int32, device=dev)
ok = True
for c in range(0, n, 1 << 28):  # chunked to bound temporaries
    x = b0[c:min(n, c + (1 << 28) + 1)] ^ flip
    ok = ok and bool((x[1:] >= x[:-1]).all())
print({"n": n, "ms": round(ms, 3), "mkeys_s": round(n / ms / 1e3), "sorted": ok, "sum_ok": int(b0.to(torch.int64).sum()) == s_in})
