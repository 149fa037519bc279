"""Reduced compute-sanitizer driver: only the schedules that rank without stability (vkrs_msd.cuh), incl. typed keys,
the device-side fallback, the moved digit window, the skipped local sort and the robust per-bucket path."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
dev = torch.device("cuda:0")
h = Handle(0, 1 << 18)
rng = np.random.default_rng(0)
def dv(a): return torch.from_numpy(a.view(np.int32)).to(dev)
cases = ((0xFFFFFFFF, 0), (0x0FFFFFFF, 0), (0xFFF, 0), (0xFFFF, 0x12340000), (0x1FFF, 0x7FFFF000))
for sched in (capi.SCHEDULE_LSD_UNSTABLE_FIRST, capi.SCHEDULE_BUCKET):
    h.set_schedule(sched)
    for n in (1, 6145, 50_001, 150_003):
        for mask, base in cases:
            k = ((rng.integers(0, 1 << 32, size=n, dtype=np.uint32) & np.uint32(mask)) + np.uint32(base)).astype(np.uint32)
            b0 = dv(k); b1 = torch.empty_like(b0)
            h.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32)); torch.cuda.synchronize()
            assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k)), (sched, n, hex(mask))
h.set_schedule(capi.SCHEDULE_BUCKET)
n = 120_001
k = ((rng.integers(0, 4, n, dtype=np.uint32) << 30) | rng.integers(0, 1 << 16, n, dtype=np.uint32)).astype(np.uint32)  # 4 big buckets: counted
b0 = dv(k); h.multi_sort(b0, torch.empty_like(b0), None, capi.multi_push_constants(n, 32)); torch.cuda.synchronize()
st = h.bucket_stats()
assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k)) and st["fallback"] == 0 and st["big_buckets"] == 4, st
m = 400_000  # half of the keys under one prefix: one big bucket, items next to it go through the redo kernel
k = np.where(rng.random(m) < 0.5, np.uint32(0x2BCD0000) | rng.integers(0, 1 << 16, m, dtype=np.uint32),
             rng.integers(0, 1 << 32, m, dtype=np.uint64).astype(np.uint32)).astype(np.uint32)
b0 = dv(k); h.multi_sort(b0, torch.empty_like(b0), None, capi.multi_push_constants(m, 32)); torch.cuda.synchronize()
st = h.bucket_stats()
assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k)) and st["fallback"] == 0 and st["big_buckets"] == 1, st
m = 1_400_000  # 300 buckets of ~4,700 keys: more than the counter pool takes -> the gated LSD passes
k = ((rng.integers(0, 300, m, dtype=np.uint32) << 16) * np.uint32(200) + rng.integers(0, 1 << 16, m, dtype=np.uint32)).astype(np.uint32)
b0 = dv(k); h.multi_sort(b0, torch.empty_like(b0), None, capi.multi_push_constants(m, 32)); torch.cuda.synchronize()
st = h.bucket_stats()
assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k)) and st["fallback"] == 1, st
for ns, mask in ((1, 0xFFFFFFFF), (100, 0xFFFFFFFF), (3000, 0xFFFF), (7000, 0xFFFFFFFF), (7164, 0), (5000, 3)):  # the one-launch small sort; its bitonic path
    k = (rng.integers(0, 1 << 32, size=ns, dtype=np.uint32) & np.uint32(mask)).astype(np.uint32)
    if mask == 3: k = (k * np.uint32(0x40000000)).astype(np.uint32)
    b0 = dv(k); b1 = torch.empty_like(b0)
    h.single_sort(b0, b1, capi.SinglePushConstants(ns)); torch.cuda.synchronize()
    assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k)), (ns, hex(mask))
k = ((rng.integers(0, 1 << 14, n, dtype=np.uint32) << 18) | rng.integers(0, 4, n, dtype=np.uint32)).astype(np.uint32)  # over-full bins: per-bucket path
b0 = dv(k); h.multi_sort(b0, torch.empty_like(b0), None, capi.multi_push_constants(n, 32)); torch.cuda.synchronize()
assert np.array_equal(b0.cpu().numpy().view(np.uint32), np.sort(k))
ints = rng.integers(-(1 << 31), 1 << 31, size=n, dtype=np.int64).astype(np.int32)
b0 = torch.from_numpy(ints.copy()).to(dev); h.multi_sort_typed(b0, torch.empty_like(b0), None, capi.multi_push_constants(n, 32), capi.KEY_I32)
torch.cuda.synchronize(); assert np.array_equal(b0.cpu().numpy(), np.sort(ints))
f = (rng.standard_normal(n) * 1e3).astype(np.float32)
b0 = torch.from_numpy(f.view(np.int32).copy()).to(dev); h.multi_sort_typed(b0, torch.empty_like(b0), None, capi.multi_push_constants(n, 32), capi.KEY_F32)
torch.cuda.synchronize(); assert np.array_equal(b0.cpu().numpy().view(np.float32), np.sort(f))
print("SANITIZE_BUCKET_PROBE_OK")
