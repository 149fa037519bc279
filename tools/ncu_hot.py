"""Summarise an `ncu --page source --csv --print-source sass` export: top stall sites.
    ncu -i X.ncu-rep --page source --csv --print-source sass > /tmp/s.csv; python tools/ncu_hot.py /tmp/s.csv [kernel_index]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# split per kernel
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
k = kernels[which]
h = k["hdr"]
ix = {n: i for i, n in enumerate(h)}
S = ix["# Samples"]
tot = sum(int(r[S] or 0) for r in k["rows"])
print(k["name"][:100], "instructions:", len(k["rows"]), "samples:", tot)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ix[n]] or 0) for r in k["rows"]) for n in stalls}
print({n: v for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
print("executed warp-insts:", sum(int(r[ix["Instructions Executed"]] or 0) for r in k["rows"]))
top = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][S] or 0))[:40]
for i in sorted(top):
    r = k["rows"][i]
    main = max(stalls, key=lambda n: int(r[ix[n]] or 0))
    print(f"{i:5d} {int(r[S]):7d} {100*int(r[S])/tot:5.1f}%  {main:22s} {r[ix['Source']][:90]}")
