#!/usr/bin/env python
"""Markdown table of an `ncu --set full` report, one row per captured launch:  python tools/ncu_table.py report.ncu-rep
Also prints the DRAM traffic per launch as JSON (for profiles/ncu_traffic.json)."""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
def f(d, k, scale=1.0, fmt="{:.1f}"):
    v = d.get(k, "")
    try: return fmt.format(float(v.replace(",", "")) * scale)
    except Exception: return "n/a"
unit = dict(zip(hdr, rows[1]))
def to_mb(d, k):
    try:
        v = float(d[k].replace(",", "")); u = unit.get(k, "")
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    except Exception: return float("nan")
def to_us(d, k):
    try:
        v = float(d[k].replace(",", "")); u = unit.get(k, "")
        return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1e-3)
    except Exception: return float("nan")
st = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
print("| kernel | duration | regs | CTAs/SM (regs / smem limit) | warps active % | warp instr | lanes/instr | issue active % | DRAM rd | DRAM wr | L2 hit % | top stalls (cycles per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
traffic = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "").split("(")[0].replace("void ", "").replace("d2gs::", "").replace("<unnamed>::", "")
    ss = sorted(((float(d[h].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                 for h in st if d.get(h) not in (None, "", "n/a")), reverse=True)[:4]
    rd, wr = to_mb(d, "dram__bytes_read.sum"), to_mb(d, "dram__bytes_write.sum")
    traffic[name] = (rd + wr) * 1e6
    print(f"| `{name}` | {to_us(d, 'gpu__time_duration.sum'):.1f} us | {f(d, 'launch__registers_per_thread', fmt='{:.0f}')} | "
          f"{f(d, 'launch__occupancy_limit_registers', fmt='{:.0f}')} / {f(d, 'launch__occupancy_limit_shared_mem', fmt='{:.0f}')} | "
          f"{f(d, 'sm__warps_active.avg.pct_of_peak_sustained_active')} | {f(d, 'smsp__inst_executed.sum', fmt='{:.3e}')} | "
          f"{f(d, 'smsp__thread_inst_executed_per_inst_executed.ratio')} | {f(d, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed')} | "
          f"{rd:.1f} MB | {wr:.1f} MB | {f(d, 'lts__t_sector_hit_rate.pct')} | " + ", ".join(f"{n} {v:.2f}" for v, n in ss) + " |")
print("\nTRAFFIC_JSON " + json.dumps(traffic))
