"""Randomised soak of the other whole-sort entry points against torch.sort: key + payload pairs (stable), uint64 keys,
vkrs_sort_auto, vkrs_single_sort, the per-stage path (vkrs_multi_sort_staged with a random tiling), for `seconds` seconds.
    python tools/fuzz_all.py [seconds] [seed]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(seed)
rng = np.random.default_rng(seed)
h = Handle(0, 1 << 20)
flip = -(1 << 31)
def rand_keys(n):
    k = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    kind = int(rng.integers(0, 4))
    if kind == 1:
        k = k & ((1 << int(rng.integers(0, 33))) - 1 if rng.integers(0, 2) else -1)
        k = k.to(torch.int32)
    elif kind == 2:
        pool = torch.randint(-(1 << 31), (1 << 31) - 1, (int(rng.integers(1, 500)),), dtype=torch.int32, device=dev, generator=g)
        k = pool[torch.randint(0, pool.numel(), (n,), device=dev, generator=g)]
    elif kind == 3:
        k = torch.sort(k)[0]
    return k.contiguous()
def fail(what, **kw):
    print(json.dumps({"fuzz_all": "MISMATCH", "what": what, "seed": seed, **kw})); sys.exit(1)
t0 = time.time(); counts = {}
while time.time() - t0 < seconds:
    what = ["pairs", "u64", "auto", "single", "staged"][int(rng.integers(0, 5))]
    if what == "pairs":
        n = int(rng.choice([rng.integers(1, 5000), rng.integers(5000, 2_000_000)]))
        k = rand_keys(n); v = torch.arange(n, dtype=torch.int32, device=dev)
        k0, v0 = k.clone(), v.clone(); k1, v1 = torch.empty_like(k), torch.empty_like(v)
        h.multi_sort_pairs(k0, k1, v0, v1, None, capi.multi_push_constants(n, 32))
        wk, idx = torch.sort(k ^ flip, stable=True)
        if not (torch.equal(k0, wk ^ flip) and torch.equal(v0.to(torch.int64), idx)): fail(what, n=n)
    elif what == "u64":
        n = int(rng.choice([rng.integers(1, 5000), rng.integers(5000, 1_000_000)]))
        k = torch.randint(-(1 << 62), (1 << 62), (n,), dtype=torch.int64, device=dev, generator=g)
        if rng.integers(0, 2): k = k & ((1 << int(rng.integers(1, 63))) - 1)
        b0, b1 = k.clone(), torch.empty_like(k)
        h.multi_sort_u64(b0, b1, None, capi.multi_push_constants(n, 32))
        f64 = -(1 << 63)
        if not torch.equal(b0, torch.sort(k ^ f64)[0] ^ f64): fail(what, n=n)
    elif what == "auto":
        n = int(rng.choice([rng.integers(1, 20000), rng.integers(20000, 5_000_000)]))
        k = rand_keys(n); b0, b1 = k.clone(), torch.empty_like(k)
        h.set_schedule(capi.SCHEDULE_AUTO)
        h.sort_auto(b0, b1, n)
        if not torch.equal(b0, torch.sort(k ^ flip)[0] ^ flip): fail(what, n=n)
    elif what == "single":
        n = int(rng.integers(1, 12289))
        k = rand_keys(n); b0, b1 = k.clone(), torch.empty_like(k)
        h.single_sort(b0, b1, capi.SinglePushConstants(n))
        if not torch.equal(b0, torch.sort(k ^ flip)[0] ^ flip): fail(what, n=n)
    else:
        n = int(rng.integers(1, 400_000)); nb = int(rng.choice([1, 2, 3, 7, 32, 100, 512]))
        k = rand_keys(n); b0, b1 = k.clone(), torch.empty_like(k)
        pc = capi.multi_push_constants(n, nb)
        hist = torch.zeros(256 * int(pc.g_num_workgroups) + 1024, dtype=torch.int32, device=dev)
        h.multi_sort_staged(b0, b1, hist, pc)
        if not torch.equal(b0, torch.sort(k ^ flip)[0] ^ flip): fail(what, n=n, nb=nb)
    h.check_device_error()
    counts[what] = counts.get(what, 0) + 1
print(json.dumps({"fuzz_all": "ok", "cases": counts, "seconds": round(time.time() - t0, 1), "seed": seed}))
