"""Fill bench_constants.json[<config>]["ncu_issue"] (and the DRAM traffic of one launch) from a committed `ncu --page raw --csv`
page of the dominant kernel, so that bench.py's roofline.issue is derived mechanically from the evidence under profiles/.
usage: python scripts/ncu_issue_constants.py <config> <profiles/xxx_raw.csv> "<what was profiled>" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg, path, what = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(path)))
hdr = rows[0]
vals = rows[2] if len(rows) > 2 else rows[1]
units = rows[1] if len(rows) > 2 else [""] * len(hdr)


def get(name, scale_units=False):
    v = float(vals[hdr.index(name)].replace(",", ""))
    if scale_units:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index(name)], 1)
    return v


issue = get("smsp__issue_active.avg.pct_of_peak_sustained_active")
lanes = get("smsp__thread_inst_executed_per_inst_executed.ratio")
out = {
    "source": os.path.relpath(path, ROOT), "what": what, "kernel": vals[hdr.index("Kernel Name")],
    "issue_active_pct": round(issue, 2), "warp_inst_per_sm_cycle": round(issue / 100 * 4, 3), "of_peak_4": round(issue / 100, 4),
    "active_lanes_per_inst": round(lanes, 2), "lane_frac": round(lanes / 32, 4),
    "useful_issue_frac": round(issue / 100 * lanes / 32, 4),
    "warps_active_pct": round(get("sm__warps_active.avg.pct_of_peak_sustained_active"), 2),
    "eligible_warps_per_cycle": round(get("smsp__warps_eligible.avg.per_cycle_active"), 3),
    "alu_pipe_pct": round(get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), 2),
    "fma_pipe_pct": round(get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2),
    "fp64_pipe_pct": round(get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 2),
    "l1_throughput_pct": round(get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), 2),
    "l1_hit_pct": round(get("l1tex__t_sector_hit_rate.pct"), 2),
    "l2_throughput_pct": round(get("lts__throughput.avg.pct_of_peak_sustained_elapsed"), 2),
    "l2_hit_pct": round(get("lts__t_sector_hit_rate.pct"), 2),
    "dram_pct": round(get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), 3),
    "registers": int(get("launch__registers_per_thread")), "launch_ms": round(get("gpu__time_duration.sum"), 3),
    "local_loads": int(get("sass__inst_executed_local_loads")), "local_stores": int(get("sass__inst_executed_local_stores")),
}
dram = int(get("dram__bytes_read.sum", True) + get("dram__bytes_write.sum", True))
p = os.path.join(ROOT, "bench_constants.json")
k = json.load(open(p))
k[cfg]["ncu_issue"] = out
k[cfg]["ncu_dram_bytes_per_launch"] = dram
k[cfg]["ncu_dram_note"] = f"dram__bytes_read.sum + dram__bytes_write.sum of one launch ({what}), {out['source']}"
json.dump(k, open(p, "w"), indent=1)
print(json.dumps(out, indent=1), dram)
