#!/usr/bin/env python3
"""profiles/ncu_r2.json from `ncu --set full` captures (run here, no GPU): per kernel the DRAM bytes per launch (`traffic` of the bench line),
L2 and L1 throughput, issue-slot utilisation, stall shares.  usage: ncu_to_json.py <kernel name>=<file.ncu-rep> ..."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = os.path.join(ROOT, "profiles", "ncu_r2.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    name, rep = arg.split("=")
    raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    hdr, val = raw[0], raw[2]
    m = {}
    for h, v in zip(hdr, val):
        try: m[h] = float(v.replace(",", ""))
        except Exception: pass
    def g(k, scale=1.0): return m.get(k) * scale if k in m else None
    units = dict(zip(raw[0], raw[1]))
    def bytes_of(k):
        if k not in m: return None
        u = units.get(k, "byte").lower()
        return m[k] * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
    dur_ms = m.get("gpu__time_duration.sum", 0) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(units.get("gpu__time_duration.sum", "ms"), 1)
    dram = (bytes_of("dram__bytes_read.sum") or 0) + (bytes_of("dram__bytes_write.sum") or 0)
    lts = m["lts__t_sectors.sum"] * 32.0 if "lts__t_sectors.sum" in m else bytes_of("lts__t_bytes.sum")
    src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout)))
    H = None; stalls = {}
    for r in src:
        if not r: continue
        if r[0] in ("Address",) or (len(r) > 1 and r[1] == "Source"): H = r; continue
        if H is None: continue
        for i, h in enumerate(H):
            if h.startswith("stall_") and "Not Issued" not in h:
                try: stalls[h] = stalls.get(h, 0) + int(r[i])
                except Exception: pass
    tot = sum(stalls.values()) or 1
    out[name] = {
        "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "") + " (ncu --set full, one launch of the 8192-sentence bench batch; numbers under ncu are not bench values)",
        "duration_ms_under_ncu": dur_ms,
        "dram_bytes_per_launch": dram,
        "dram_gbs": dram / (dur_ms / 1e3) / 1e9 if dur_ms else None,
        "l2": {"bytes_per_launch": lts, "gbs": lts / (dur_ms / 1e3) / 1e9 if (lts and dur_ms) else None, "sector_hit_rate_pct": m.get("lts__t_sector_hit_rate.pct"), "pct_of_peak": m.get("lts__t_sectors.sum.pct_of_peak_sustained_elapsed")},
        "l1": {"sector_hit_rate_pct": m.get("l1tex__t_sector_hit_rate.pct")},
        "issue": {"issue_active_pct": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "warps_active_pct": m.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                  "warp_instructions": m.get("smsp__inst_executed.sum"), "threads_per_instruction": m.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                  "registers_per_thread": m.get("launch__registers_per_thread")},
        "tensor_pipe_active_pct": m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "local_loads": m.get("sass__inst_executed_local_loads"), "local_stores": m.get("sass__inst_executed_local_stores"),
        "stalls_pct": {k.replace("stall_", ""): round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]},
    }
    print(name, json.dumps(out[name])[:600])
json.dump(out, open(out_path, "w"), indent=1)
