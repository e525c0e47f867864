#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw metrics, stall sampling, instruction share per function) — run here, no GPU needed."""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, unit, val = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__sass_inst_executed_op_shared_ld.sum", "l1tex__t_bytes.sum", "lts__t_sectors_op_read.sum"]
for h, u, v in zip(hdr, unit, val):
    if h in want: print("%-60s %-12s %s" % (h, u, v))
src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda,sass"]))))
cur = None; H = None; data = []; stalls = {}
for r in src:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; continue
    if H is None: continue
    if r[0] != "":
        try:
            data.append((int(r[H.index("# Samples")]), int(r[H.index("Instructions Executed")]), int(r[H.index("Thread Instructions Executed")]), cur, int(r[0]), r[1]))
        except Exception: pass
    else:
        for i, h in enumerate(H):
            if h.startswith("stall_") and "Not Issued" not in h:
                try: stalls[h] = stalls.get(h, 0) + int(r[i])
                except Exception: pass
ts = sum(d[0] for d in data); ti = sum(d[1] for d in data)
print("total samples", ts, "total warp instructions", ti)
print("stalls:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, sum(stalls.values()))) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]))
# function regions from the source listing
listing = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--print-source", "cuda"]))))
files = {}; cf = None
for r in listing:
    if r and r[0] == "File Name": cf = r[1].split("/")[-1]; files[cf] = {}
    elif len(r) >= 2 and r[0].isdigit() and cf: files[cf][int(r[0])] = r[1]
marks = {}
for f, lines in files.items():
    ms = []
    for ln, s in sorted(lines.items()):
        m = re.search(r"__(?:device|global)__.*?\b(\w+)\s*\(", s)
        if m and ("{" not in s or s.strip().endswith("{") or True) and not s.strip().startswith("//"):
            if re.search(r"\b(if|for|while|return)\b", s.split("(")[0]) is None: ms.append((ln, m.group(1)))
    marks[f] = ms
def region(f, l):
    r = f + ":top"
    for ln, n in marks.get(f, []):
        if l >= ln: r = f + ":" + n
    return r
agg = {}
for s, i, t, f, l, _ in data:
    a = agg.setdefault(region(f, l), [0, 0, 0]); a[0] += s; a[1] += i; a[2] += t
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:22]:
    print("%-44s samples %5.1f%%  inst %5.1f%% (%.3fG)  avg threads %.1f" % (k, 100 * a[0] / max(1, ts), 100 * a[1] / max(1, ti), a[1] / 1e9, a[2] / max(1, a[1])))
print("hottest lines:")
for d in sorted(data, reverse=True)[:25]:
    print("%5.1f%% inst=%10d %s:%d %s" % (100 * d[0] / ts, d[1], d[3], d[4], d[5].strip()[:110]))
