"""Stage-by-stage check of one bucket-schedule sort (bucket_probe.diagnose).  python tools/diag_case.py <dist> <n>"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv, args = sys.argv[:1], sys.argv[1:]
import importlib.util
spec = importlib.util.spec_from_file_location("bp", os.path.join(os.path.dirname(os.path.abspath(__file__)), "bucket_probe.py"))
bp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bp)
from vkradixsort_b200 import Handle, capi
dist, n = args[0], int(float(args[1]))
h = Handle(0, n)
h.set_schedule(capi.SCHEDULE_BUCKET)
keys = bp.gen(dist, n, 1000 + n)
want = bp.expect_sorted(keys)
rep = {}
for stage in (1, 2):
    b0, b1 = keys.clone(), torch.full_like(keys, 0x5A5A5A5A)
    h.debug_bucket_stop(stage)
    h.multi_sort(b0, b1, None, capi.multi_push_constants(n, 32))
    torch.cuda.synchronize()
    st = h.bucket_stats()
    res = b1 if stage == 1 else b0
    base = st["key_min"] if st["recount"] else 0
    perm_ok = bool(torch.equal(bp.expect_sorted(res), want))
    sh = st["shift1"] if stage == 1 else st["shift2"]
    d = (((res - base) >> sh) & (255 if stage == 1 else 0xFFFF)).to(torch.int64)
    grouped = bool((d[1:] >= d[:-1]).all())
    nbad = int((d[1:] < d[:-1]).sum())
    print(json.dumps({"stage": stage, "permutation": perm_ok, "grouped": grouped, "descents": nbad, "stats": st}), flush=True)
