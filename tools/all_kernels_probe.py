"""Runs every kernel of the library once on its natural size (for an ncu launch list with DRAM bytes)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
dev = torch.device("cuda:0")
n = 100_000_000
h = Handle(0, n)
keys = torch.from_numpy(np.random.default_rng(1).integers(0, 1 << 32, size=n, dtype=np.uint32).view(np.int32)).to(dev)
b0, b1 = keys.clone(), torch.empty_like(keys)
pc = capi.multi_push_constants(n, 32)
for v in (0, 4, 6):  # segmented (default), pipelined one-sweep, simple one-sweep
    h.set_variant(v); b0.copy_(keys); h.multi_sort(b0, b1, None, pc)
h.set_variant(0)
vals = torch.arange(n, dtype=torch.int32, device=dev); v1 = torch.empty_like(vals)
b0.copy_(keys); h.multi_sort_pairs(b0, b1, vals, v1, None, pc)
k64 = torch.randint(0, 1 << 62, (n // 2,), dtype=torch.int64, device=dev); k64b = torch.empty_like(k64)
h.multi_sort_u64(k64, k64b, None, capi.multi_push_constants(n // 2, 32))
pcs = capi.multi_push_constants(n, 4096)   # the reference's best nb at 10^8
hist = torch.zeros(pcs.g_num_workgroups * 256, dtype=torch.int32, device=dev)
b0.copy_(keys); h.multi_sort_staged(b0, b1, hist, pcs)
small = keys[:10_000].clone(); h.single_sort(small, torch.empty_like(small), capi.SinglePushConstants(10_000))
mm = torch.zeros(2, dtype=torch.int32, device=dev); cnt = torch.zeros(256, dtype=torch.int32, device=dev)
h.key_range(keys, n, mm); h.partition(keys, b1, n, 0, 24, cnt)
torch.cuda.synchronize(); h.check_device_error(); print("ALL_KERNELS_OK")
