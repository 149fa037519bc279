"""Small driver for compute-sanitizer runs: every whole-sort entry point once on modest sizes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
dev = torch.device("cuda:0")
h = Handle(0, 1 << 18)
rng = np.random.default_rng(0)
def dv(a): return torch.from_numpy(a.view(np.int32 if a.dtype == np.uint32 else np.int64)).to(dev)
for n in (1, 1000, 50_001, 200_003):
    k = rng.integers(0, 1 << 32, size=n, dtype=np.uint32)
    b0 = dv(k); b1 = torch.empty_like(b0)
    h.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32)); torch.cuda.synchronize()
    assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k))
    v = np.arange(n, dtype=np.uint32); k0, v0 = dv(k), dv(v)
    h.multi_sort_pairs(k0, torch.empty_like(k0), v0, torch.empty_like(v0), None, capi.multi_push_constants(n, 32)); torch.cuda.synchronize()
    assert np.array_equal(v0.cpu().numpy().view(np.uint32), np.argsort(k, kind="stable").astype(np.uint32))
    k64 = rng.integers(0, 1 << 63, size=n, dtype=np.uint64); q0 = dv(k64)
    h.multi_sort_u64(q0, torch.empty_like(q0), None, capi.multi_push_constants(n, 32)); torch.cuda.synchronize()
    assert np.array_equal(q0.cpu().numpy().view(np.uint64), np.sort(k64))
    b0 = dv(k); hist = torch.zeros(capi.multi_push_constants(n, 3).g_num_workgroups * 256, dtype=torch.int32, device=dev)
    h.multi_sort_staged(b0, torch.empty_like(b0), hist, capi.multi_push_constants(n, 3)); torch.cuda.synchronize()
    assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k))
    if n <= 50_001:
        b0 = dv(k); h.single_sort(b0, torch.empty_like(b0), capi.SinglePushConstants(n)); torch.cuda.synchronize()
        assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k))
    mm = torch.zeros(2, dtype=torch.int32, device=dev); cnt = torch.zeros(256, dtype=torch.int32, device=dev)
    h.key_range(dv(k), n, mm); h.partition(dv(k), torch.empty(n, dtype=torch.int32, device=dev), n, 0, 24, cnt); torch.cuda.synchronize()
# the schedules that rank without stability (vkrs_msd.cuh): uniform, 28-bit (moved digit window + recount),
# 12-bit (no local sort) and one-bucket keys (device-side fallback to the LSD passes)
for sched in (capi.SCHEDULE_LSD_UNSTABLE_FIRST, capi.SCHEDULE_BUCKET):
    h.set_schedule(sched)
    for n in (1, 1000, 6145, 50_001, 200_003):
        for mask, base in ((0xFFFFFFFF, 0), (0x0FFFFFFF, 0), (0xFFF, 0), (0xFFFF, 0x12340000)):
            k = (rng.integers(0, 1 << 32, size=n, dtype=np.uint32) & np.uint32(mask)) | np.uint32(base)
            b0 = dv(k); b1 = torch.empty_like(b0)
            h.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32)); torch.cuda.synchronize()
            assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k)), (sched, n, hex(mask))
h.set_schedule(capi.SCHEDULE_AUTO)
print("SANITIZE_PROBE_OK")
