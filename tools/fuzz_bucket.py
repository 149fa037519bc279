"""Randomised soak of the keys-only schedules against torch.sort: random sizes, bit masks, offsets, duplicate levels,
hot prefixes, alignments (sub-buffers that start 4 / 8 / 12 bytes into an allocation), schedules and key types, for
`seconds` seconds.  Prints one JSON line; exits 1 at the first mismatch.
    python tools/fuzz_bucket.py [seconds] [seed]"""
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
scheds = [capi.SCHEDULE_BUCKET, capi.SCHEDULE_BUCKET, capi.SCHEDULE_AUTO, capi.SCHEDULE_LSD_UNSTABLE_FIRST, capi.SCHEDULE_LSD]
flip = -(1 << 31)
t0 = time.time(); cases = 0; keys_total = 0; kinds = {}
while time.time() - t0 < seconds:
    n = int(rng.choice([rng.integers(1, 9000), rng.integers(9000, 300_000), rng.integers(300_000, 6_000_000)]))
    k = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev, generator=g)
    kind = int(rng.integers(0, 7))
    if kind == 1:   # masked bits + offset (narrow ranges, ranges across 2^31)
        bits = int(rng.integers(0, 33)); off = int(rng.integers(-(1 << 31), 1 << 31))
        k = (k & ((1 << bits) - 1 if bits < 32 else -1)) + off
        k = ((k + (1 << 31)) % (1 << 32) - (1 << 31)).to(torch.int32)
    elif kind == 2:  # few distinct values
        pool = torch.randint(-(1 << 31), (1 << 31) - 1, (int(rng.integers(1, 3000)),), dtype=torch.int32, device=dev, generator=g)
        k = pool[torch.randint(0, pool.numel(), (n,), device=dev, generator=g)]
    elif kind == 3:  # a hot 16-bit prefix
        hot = torch.rand(n, device=dev, generator=g) < float(rng.uniform(0.05, 0.95))
        k = torch.where(hot, (k & 0xFFFF) | (int(rng.integers(0, 1 << 15)) << 16), k)
    elif kind == 4:  # sorted / reversed
        k = torch.sort(k, descending=bool(rng.integers(0, 2)))[0]
    elif kind == 5:  # clusters
        c = torch.randint(-(1 << 31), (1 << 31) - 1, (int(rng.integers(1, 40)),), dtype=torch.int64, device=dev, generator=g)
        k = (c[torch.randint(0, c.numel(), (n,), device=dev, generator=g)] + (torch.randn(n, device=dev, generator=g) * float(10 ** rng.uniform(0, 7))).to(torch.int64))
        k = ((k + (1 << 31)) % (1 << 32) - (1 << 31)).to(torch.int32)
    off = int(rng.integers(0, 4))
    big0 = torch.zeros(n + 8, dtype=torch.int32, device=dev); big1 = torch.zeros(n + 8, dtype=torch.int32, device=dev)
    big0[off:off + n] = k
    sched = scheds[int(rng.integers(0, len(scheds)))]
    h.set_schedule(sched)
    typed = int(rng.integers(0, 3))
    pc = capi.multi_push_constants(n, 32)
    if typed == 0:
        h.multi_sort(big0[off:], big1[off:], None, pc)
        want = (torch.sort(k ^ flip)[0]) ^ flip          # unsigned order
    elif typed == 1:
        h.multi_sort_typed(big0[off:], big1[off:], None, pc, capi.KEY_I32)
        want = torch.sort(k)[0]
    else:
        f = k.view(torch.float32)
        finite = ~torch.isnan(f)
        k = torch.where(finite, k, torch.zeros_like(k)); big0[off:off + n] = k   # no NaNs: their order is a convention
        h.multi_sort_typed(big0[off:], big1[off:], None, pc, capi.KEY_F32)
        bits = k.to(torch.int64) & 0xFFFFFFFF
        key = torch.where(bits >= (1 << 31), (~bits) & 0xFFFFFFFF, bits | (1 << 31))   # the order-preserving map
        want = k[torch.sort(key, stable=True)[1]]
    h.check_device_error()
    got = big0[off:off + n]
    ok = bool(torch.equal(got, want)) and int(big0[:off].abs().sum()) == 0 and int(big0[off + n:].abs().sum()) == 0
    if not ok:
        bad = torch.nonzero(got != want).flatten()
        i = int(bad[0]) if bad.numel() else -1
        info = {"fuzz": "MISMATCH", "case": cases, "n": n, "kind": kind, "sched": capi.schedule_name(sched), "typed": typed, "off": off, "seed": seed,
                "mismatches": int(bad.numel()), "first": i, "last": int(bad[-1]) if bad.numel() else -1,
                "got": [hex(int(x) & 0xFFFFFFFF) for x in got[max(0, i - 2):i + 4].tolist()],
                "want": [hex(int(x) & 0xFFFFFFFF) for x in want[max(0, i - 2):i + 4].tolist()],
                "same_multiset": bool(torch.equal(torch.sort(got)[0], torch.sort(want)[0])),
                "stats": h.bucket_stats()}
        # which combinations of schedule and key type get these keys wrong?
        combos = {}
        for sc in (capi.SCHEDULE_BUCKET, capi.SCHEDULE_LSD_UNSTABLE_FIRST, capi.SCHEDULE_LSD):
            for ty, kt in ((0, None), (1, capi.KEY_I32), (2, capi.KEY_F32)):
                a = k.clone(); b = torch.empty_like(a)
                h.set_schedule(sc)
                if kt is None:
                    h.multi_sort(a, b, None, pc); w = (torch.sort(k ^ flip)[0]) ^ flip
                else:
                    h.multi_sort_typed(a, b, None, pc, kt)
                    w = torch.sort(k)[0] if ty == 1 else want
                combos[f"{capi.schedule_name(sc)[:6]}/{ty}"] = bool(torch.equal(a, w))
        info["combos_ok"] = combos
        torch.save(k.cpu(), "gpurun_out/fuzz_fail_keys.pt")
        print(json.dumps(info))
        sys.exit(1)
    cases += 1; keys_total += n; kinds[kind] = kinds.get(kind, 0) + 1
print(json.dumps({"fuzz": "ok", "cases": cases, "keys": keys_total, "seconds": round(time.time() - t0, 1), "kinds": kinds, "seed": seed}))
