"""Per-kernel instruction census of the shipped library: which sm_100a mechanisms each kernel really uses.
    python tools/sass_census.py [lib.so] > profiles/r02_sass_census.txt
Reads `cuobjdump -sass` (mnemonics) and `cuobjdump -res-usage` (registers, shared memory).  UBLKCP = 1-D TMA bulk copy
(cp.async.bulk), SYNCS = mbarrier operations, LDGSTS = cp.async, ATOMS / RED = shared / global atomics, VOTE / SHFL /
MATCH = warp-level exchange, BAR = block or named barriers."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vkradixsort_b200", "lib", "libvkradixsort_b200.so")
COLS = ["UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "ATOMG", "RED", "VOTE", "SHFL", "MATCH", "BAR", "LDS", "STS", "LDG", "STG", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, total, arch, cur = {}, {}, set(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for c in COLS:
            if op == c or op.startswith(c + ".") or (c in ("ATOMG",) and op == "ATOM") or (c == "BAR" and op in ("BAR", "WARPSYNC")):
                counts[cur][c] += 1
            elif op.startswith(c) and c in ("UBLKCP", "SYNCS", "LDGSTS", "VOTE", "SHFL", "MATCH", "RED"):
                counts[cur][c] += 1
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage, fn = {}, None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*SHARED:(\d+)", line)
    if m and fn:
        usage[fn] = (int(m.group(1)), int(m.group(2)))
names = demangle(list(counts))
print(f"# SASS census of {os.path.relpath(lib, ROOT)} -- cubin architectures: {sorted(arch)}")
print("# static shared memory only (dynamic shared memory is requested at launch: see DESIGN.md section 4)")
print(f"{'kernel':72s} {'inst':>6s} {'regs':>4s} {'smem':>6s} " + " ".join(f"{c:>6s}" for c in COLS))
for k in sorted(counts, key=lambda k: names[k]):
    short = re.sub(r"\(.*", "", names[k]).replace("vkrs::", "").replace("void ", "")
    short = re.sub(r"\(int\)|\(bool\)|unsigned ", "", short)
    reg, smem = usage.get(k, (0, 0))
    print(f"{short[:72]:72s} {total[k]:6d} {reg:4d} {smem:6d} " + " ".join(f"{counts[k][c]:6d}" for c in COLS))
