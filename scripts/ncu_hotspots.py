"""Summarise an .ncu-rep: headline metrics + per-source-line instruction share and average active lanes.
usage: python scripts/ncu_hotspots.py report.ncu-rep [out.txt] [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed.sum", "smsp__inst_executed.sum"]
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if h in WANT:
            out.write(f"{h:70s} {units[i]:16s} {r[i]}\n")
    out.write("\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
h2 = None
lines = {}
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h2 = r
        continue
    if r[0] not in ("", "Function Name") and h2:
        d = dict(zip(h2, r))
        try:
            key = (cur, int(r[0]))
            v = (float(d["Instructions Executed"] or 0), float(d["Thread Instructions Executed"] or 0), float(d["# Samples"] or 0), r[1].strip()[:100])
        except (ValueError, KeyError):
            continue
        old = lines.get(key, (0, 0, 0, v[3]))
        lines[key] = (old[0] + v[0], old[1] + v[1], old[2] + v[2], v[3])
ti = sum(v[0] for v in lines.values()) or 1
tt = sum(v[1] for v in lines.values())
ts = sum(v[2] for v in lines.values()) or 1
out.write(f"warp instructions {ti:.4g}, thread instructions {tt:.4g}, AVERAGE ACTIVE LANES PER INSTRUCTION {tt / ti:.2f} of 32\n")
pf = collections.defaultdict(lambda: [0, 0, 0])
for (f, _), v in lines.items():
    pf[f][0] += v[0]; pf[f][1] += v[1]; pf[f][2] += v[2]
for f, v in sorted(pf.items(), key=lambda kv: -kv[1][0]):
    out.write(f"  {f:28s} inst {100 * v[0] / ti:5.1f}%  samples {100 * v[2] / ts:5.1f}%  avg active {v[1] / max(v[0], 1):5.2f}\n")
out.write("\nfile:line  inst%  samples%  avg-active-lanes | source\n")
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    out.write(f"{f}:{l}  {100 * v[0] / ti:.2f}  {100 * v[2] / ts:.2f}  {v[1] / max(v[0], 1):.1f} | {v[3]}\n")
