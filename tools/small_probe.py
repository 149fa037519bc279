"""Small sorts: CUDA-event time of one vkrs_single_sort call (n <= 7676: small_sort_kernel, one launch) next to the event
time of an empty stream (the floor of this way of timing).  Under `ncu --metrics gpu__time_duration.sum` the launch
list gives the kernels' own durations."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vkradixsort_b200 import Handle, capi
dev = torch.device("cuda:0")
h = Handle(0, 1 << 16)
def timed(fn, reps=200):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return round(1e3 * t[len(t) // 2], 2)
print(json.dumps({"empty_event_pair_us": timed(lambda: None)}))
for n in (100, 1000, 2048, 4096, 7676, 7677, 12288):
    k = torch.randint(-(1 << 31), (1 << 31) - 1, (n,), dtype=torch.int32, device=dev)
    b0, b1 = k.clone(), torch.empty_like(k)
    pc = capi.SinglePushConstants(n)
    for _ in range(5): h.single_sort(b0, b1, pc)
    print(json.dumps({"n": n, "single_sort_us": timed(lambda: h.single_sort(b0, b1, pc))}))
