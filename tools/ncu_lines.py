#!/usr/bin/env python
"""summarise an .ncu-rep: headline metrics + per-source-line instruction / stall-sample shares (needs -lineinfo)
usage: tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, rows = r[0], r[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
for i, n in enumerate(h):
    if n in want:
        print(f"{n:70s}", [row[i] for row in rows])
for i, n in enumerate(h):
    if "smsp__average_warps_issue_stalled" in n and "per_issue_active" in n and "not_issued" not in n:
        v = [float(row[i] or 0) for row in rows]
        if max(v) > 0.3:
            print(f"{n:90s}", v)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None; agg = []; cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10 or r[0] == "": continue
    iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    try:
        agg.append((int(r[iI]), int(r[iS]), int(r[iT]), cur, r[0], r[1][:100]))
    except ValueError:
        continue
tot = sum(a[0] for a in agg) or 1; tots = sum(a[1] for a in agg) or 1
print("total warp-inst", tot, "samples", tots)
for a in sorted(agg, reverse=True)[:top]:
    print(f"{a[0]/tot*100:5.1f}% inst {a[1]/tots*100:5.1f}% samp thr/inst {a[2]/max(a[0],1):5.1f}  {a[3]}:{a[4]}  {a[5]}")
