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
    f.write(f"# ncu launch list, {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400` around\n"
            "`python bench.py --steps 2 --warmup 3 --md-steps 5 --equil 30 --no-cpu-baseline --no-e2e` (C2, N = 240 000).\n"
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
