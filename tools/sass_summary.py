#!/usr/bin/env python
"""SASS mnemonic census of libd2gs.so (sm_100a): python tools/sass_summary.py > profiles/sass_<round>.md
Counts, per kernel, the instructions that show which hardware paths the kernels use: bulk async copies (UBLKCP = TMA 1-D
bulk, LDGSTS = cp.async), mbarrier operations (SYNCS), global reductions (RED/REDG, incl. 16-byte F32x4), warp votes /
matches / shuffles, MUFU, and the absence of tensor-core (HMMA / UTC*MMA) and tensor-memory (LDTM / UTMALDG) instructions."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dynamic-2dgs_b200", "d2gs_b200", "libd2gs.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UBLKCP", "LDGSTS", "SYNCS", "REDG", "F32x4", "ATOMG", "ATOMS", "VOTE", "MATCH", "SHFL", "REDUX", "MUFU", "LDS", "STS", "LDG", "STG",
        "BAR", "HMMA", "UTC", "LDTM", "UTMALDG", "STL", "LDL"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
    if cur and m:
        op = m.group(1)
        per[cur]["total"] += 1
        for k in keys:
            if k in op: per[cur][k] += 1
arch = re.findall(r"arch = (sm_\w+)", out)
print(f"# SASS census of `{os.path.relpath(lib, ROOT)}` ({', '.join(sorted(set(arch)))}; cuobjdump -sass)\n")
print("| kernel | instr | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
def demangle(n):
    try:
        d = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
        i = d.find("(")
        return d if i < 0 else d[:i]
    except Exception: return n
tot = collections.Counter()
for f, c in per.items():
    if c["total"] < 40: continue
    name = demangle(f).replace("d2gs::", "").replace("(anonymous namespace)::", "").replace("void ", "")[:48]
    print(f"| `{name}` | {c['total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in keys) + " |")
    tot.update(c)
print(f"| **all kernels** | {tot['total']} | " + " | ".join(str(tot[k]) if tot[k] else "0" for k in keys) + " |")
