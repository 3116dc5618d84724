#!/usr/bin/env python
"""gpurun_out/launches_<tag>.csv + gpurun_out/pair_<tag>.ncu-rep -> profiles/<tag>_launches.md, profiles/<tag>_pair_kernel.md,
profiles/<tag>_pair_kernel.json (read by bench.py for roofline.traffic).  usage: tools/summarize_profiles.py <tag>"""
import collections, csv, io, json, os, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles"); os.makedirs(out, exist_ok=True)
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv"))))
for k, r in enumerate(rows):
    if "Kernel Name" in r:
        h, data = r, rows[k + 1:]; break
iN, iV, iM = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
agg = collections.OrderedDict()
for r in data:
    if len(r) > iV and r[iM] == "gpu__time_duration.sum":
        agg.setdefault(r[iN].split("(")[0].replace("void ", "").replace("smd::", ""), []).append(float(r[iV].replace(",", "")) / 1e3)
tot = sum(sum(v) for k, v in agg.items() if not k.startswith("at::"))
with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
    f.write(f"# ncu launch list, {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500` around\n"
            "`python bench.py --steps 2 --warmup 3 --md-steps 8 --equil 32 --no-cpu-baseline --no-e2e` (C2, N = 240 000; tools/profile_round.sh).\n"
            "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live phase timing, not absolutes.\n\n"
            "| kernel | launches | avg us | share of our kernels |\n|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        if k.startswith("at::"):
            continue
        f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.2f} | {100*sum(v)/tot:.1f} % |\n")
rep = os.path.join(ROOT, "gpurun_out", f"pair_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw))); h, row = r[0], r[2]
g = lambda n: row[h.index(n)] if n in h else None
units = r[1]
def bytes_of(n):
    v, u = float(g(n)), units[h.index(n)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
traffic = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
json.dump({"kernel": g("Kernel Name").split("(")[0], "dram_bytes_per_launch": traffic, "duration_us_under_ncu": float(g("gpu__time_duration.sum")),
           "source": f"ncu --set full, profiles/{tag}_pair_kernel.md"}, open(os.path.join(out, f"{tag}_pair_kernel.json"), "w"))
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, "25"], capture_output=True, text=True).stdout
with open(os.path.join(out, f"{tag}_pair_kernel.md"), "w") as f:
    f.write(f"# ncu --set full, pair force kernel, {tag}\n\n`{g('Kernel Name')[:80]}...` one launch, C2 (N = 240 000), `--clock-control none`.\n\n| metric | value | unit |\n|---|---|---|\n")
    for k in keys:
        if k in h:
            f.write(f"| {k} | {g(k)} | {units[h.index(k)]} |\n")
    f.write(f"\nDRAM traffic per launch = {traffic/1e6:.2f} MB (algorithmic: 52 B x 240 000 = 12.48 MB).\n\n## hottest source lines (instruction share, stall-sample share, active threads per instruction)\n\n```\n")
    f.write("\n".join(l for l in lines.splitlines() if "% inst" in l or l.startswith("total")))
    f.write("\n```\n")
print("wrote", os.listdir(out))

# ---- the other kernels of the step (cell build, fused integrator seam, energy kernel): HBM view
rep2 = os.path.join(ROOT, "gpurun_out", f"others_{tag}.ncu-rep")
if os.path.exists(rep2):
    raw = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw))); h, units, rows = r[0], r[1], r[2:]
    def col(row, n):
        return row[h.index(n)] if n in h else ""
    def byt(row, n):
        v, u = float(col(row, n) or 0), units[h.index(n)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6456.2)
    seen = {}
    for row in rows:
        name = col(row, "Kernel Name").split("(")[0].replace("void ", "").replace("smd::", "")
        if name in seen:
            continue
        us = float(col(row, "gpu__time_duration.sum"))
        us *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[h.index("gpu__time_duration.sum")], 1)
        traffic = byt(row, "dram__bytes_read.sum") + byt(row, "dram__bytes_write.sum")
        seen[name] = (us, traffic, col(row, "launch__registers_per_thread"), col(row, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                      col(row, "lts__t_sector_hit_rate.pct"), col(row, "smsp__issue_active.avg.pct_of_peak_sustained_active"))
    with open(os.path.join(out, f"{tag}_other_kernels.md"), "w") as f:
        f.write(f"# ncu --set full, the other kernels of one MD step, {tag}\n\nC2 (N = 240 000), one launch each, `--clock-control none`; cold-cache, serialised launches "
                f"(durations are upper bounds of the in-step ones). HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json).\n\n"
                "| kernel | us | DRAM MB (read + write) | achieved GB/s | of HBM peak | regs | warps active % | L2 hit % | issue active % |\n|---|---|---|---|---|---|---|---|---|\n")
        for name, (us, tr, regs, wa, l2, ia) in seen.items():
            gbs = tr / (us * 1e-6) / 1e9 if us > 0 else 0
            f.write(f"| `{name}` | {us:.2f} | {tr/1e6:.2f} | {gbs:.0f} | {100*gbs/peak:.1f} % | {regs} | {float(wa or 0):.1f} | {float(l2 or 0):.1f} | {float(ia or 0):.1f} |\n")
        f.write("\nC2's whole state (240 000 x ~290 B = 70 MB) stays resident in the 126 MB L2 from step to step, so these kernels hardly touch DRAM: "
                "they are bound by launch + dependent-load latency at this size, not by HBM (see DESIGN.md section 3).\n")
    print("wrote", f"{tag}_other_kernels.md")
