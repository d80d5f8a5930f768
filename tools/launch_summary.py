#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-kernel table of ONE step
(the launches between the last two `mlp_transpose_kernel` launches):  python tools/launch_summary.py launches.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [(r[kn], float(r[mv].replace(",", "")) / 1000.0) for r in data if len(r) > mv]
idx = [i for i, n in enumerate(names) if "mlp_transpose" in n[0]]
step = names[idx[-2]:idx[-1]] if len(idx) >= 2 else names
agg = collections.OrderedDict()
for n, us in step:
    short = n.split("(")[0].replace("void ", "")[:70]
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
own = sum(v[1] for k, v in agg.items() if "d2gs::" in k or "cub::" in k)
nown = sum(v[0] for k, v in agg.items() if "d2gs::" in k or "cub::" in k)
print(f"launches per step: **{len(step)}**, kernel time {tot:.0f} us, of which own kernels + CUB {own:.0f} us ({100 * own / tot:.1f} %) in {nown} launches, "
      f"torch glue {tot - own:.0f} us in {len(step) - nown} launches\n")
print("| kernel | launches | us | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")
