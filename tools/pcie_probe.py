"""Host <-> device copy bandwidth of a 400 MB pinned buffer: one copy vs 2 / 4 concurrent chunks on separate streams.
    python tools/pcie_probe.py"""
import json, time
import torch
n = 100_000_000
dev = torch.device("cuda:0")
host = torch.empty(n, dtype=torch.int32).pin_memory()
host.random_()
d = torch.empty(n, dtype=torch.int32, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]
def run(direction, chunks):
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        per = n // chunks
        for c in range(chunks):
            with torch.cuda.stream(streams[c]):
                sl = slice(c * per, n if c == chunks - 1 else (c + 1) * per)
                if direction == "h2d":
                    d[sl].copy_(host[sl], non_blocking=True)
                else:
                    host[sl].copy_(d[sl], non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return 4 * n / ts[len(ts) // 2] / 1e9
for direction in ("h2d", "d2h"):
    print(json.dumps({"direction": direction, **{f"gbs_{c}_chunks": round(run(direction, c), 1) for c in (1, 2, 4)}}), flush=True)
# both directions at once (full duplex)
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(streams[0]): d[: n // 2].copy_(host[: n // 2], non_blocking=True)
with torch.cuda.stream(streams[1]): host[n // 2:].copy_(d[n // 2:], non_blocking=True)
torch.cuda.synchronize(); print(json.dumps({"duplex_200MB_each_ms": round(1e3 * (time.perf_counter() - t0), 2)}))
