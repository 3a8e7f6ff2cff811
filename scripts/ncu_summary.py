#!/usr/bin/env python
"""Summarise an ncu report: per-opcode executed instructions, stall reasons, DRAM bytes (run locally)."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]; npx = float(sys.argv[2]) if len(sys.argv) > 2 else 119771136.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
si, ie, ws = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ops, samp, inst = collections.Counter(), collections.Counter(), []
for r in rows[hi + 1:]:
    if len(r) <= max(si, ie, ws) or not r[ie].isdigit():
        continue
    n, s_, text = int(r[ie]), int(r[ws]), r[si].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", text)
    base = (m.group(2) if m else text).split(".")[0]
    ops[base] += n; samp[base] += s_; inst.append((n, s_, text))
tot, stot, px = sum(ops.values()), max(1, sum(samp.values())), npx / 32
print(f"warp instructions {tot}  = {tot / px:.2f} per pixel")
for k, v in ops.most_common(16):
    print(f"  {k:<8} {v / px:6.2f}/px   stall samples {100 * samp[k] / stot:5.1f}%")
mx = max(n for n, _, _ in inst)
print("static instructions", len(inst), " hot (>= max/4):", sum(1 for n, _, _ in inst if n >= mx / 4))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
keys = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
keys += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        v = [r[i] for r in data]
        try:
            if all(float(x) < 0.05 for x in v) and "stalled" in k:
                continue
        except ValueError:
            pass
        print(f"  {k.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')}: {v}")
