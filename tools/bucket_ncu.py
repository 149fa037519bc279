"""Workload for ncu captures of one keys-only schedule: a few whole sorts of n uniform keys.
    ncu --set full --clock-control none --import-source on -k regex:'msd_(scatter|local)' -s 3 -c 3 -o gpurun_out/x \
        python tools/bucket_ncu.py [n] [schedule] [sorts] [hot_prefix]
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
sched = int(sys.argv[2]) if len(sys.argv) > 2 else capi.SCHEDULE_BUCKET
sorts = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(5)
keys = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
if len(sys.argv) > 4 and sys.argv[4] == "hot_prefix":  # half of the keys under one 16-bit prefix (big-bucket kernels)
    hot = torch.rand(n, device=dev, generator=g) < 0.5
    keys = torch.where(hot, (keys & 0xFFFF) | 0x2BCD0000, keys)
b0, b1 = torch.empty_like(keys), torch.empty_like(keys)
h = Handle(0, n)
h.set_schedule(sched)
for _ in range(sorts):
    b0.copy_(keys)
    h.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32))
torch.cuda.synchronize()
print("ok", bool(((b0[1:] ^ -(1 << 31)) >= (b0[:-1] ^ -(1 << 31))).all()))
