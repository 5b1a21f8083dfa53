"""Turns an ncu capture (gpurun_out/*.ncu-rep) + launch list into the committed summaries under profiles/.
usage: summarize_profile.py <solve.ncu-rep> <launches.csv | -> <tag> [--not-latest]   ("-": no launch list)"""
import csv, json, os, subprocess, sys, collections, re

rep, launches, tag = sys.argv[1:4]
latest = "--not-latest" not in sys.argv
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda k: vals[hdr.index(k)] if k in hdr else None
unit = lambda k: units[hdr.index(k)] if k in hdr else None
def num(k):
    v = get(k)
    if v is None: return None
    x = float(v.replace(",", ""))
    u = unit(k) or ""
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1)
    return x * scale
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active", "sm__icc_request_hit_rate.pct",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
summ = {k: num(k) for k in keys}
summ["kernel"] = rows[2][hdr.index("Kernel Name")] if "Kernel Name" in hdr else None
dram = (summ["dram__bytes_read.sum"] or 0) + (summ["dram__bytes_write.sum"] or 0)
out = {"tag": tag, "solve_kernel": {"dram_bytes_per_launch": dram, "duration_s_under_ncu": summ["gpu__time_duration.sum"], "metrics": summ}}
# launch list: per-kernel totals and shares
tot = collections.defaultdict(lambda: [0, 0.0])
for r in (csv.reader(open(launches)) if launches != "-" else []):
    if len(r) > 14 and r[12] == "gpu__time_duration.sum":
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
        tot[name][0] += 1
        tot[name][1] += float(r[14]) * 1e-9
allt = sum(v[1] for v in tot.values()) or 1.0
out["launch_list"] = {k: {"launches": v[0], "seconds": v[1], "share": v[1] / allt} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", f"ncu_summary_{tag}.json"), "w") as f:
    json.dump(out, f, indent=1)
if latest:
    with open(os.path.join(ROOT, "profiles", "ncu_summary_latest.json"), "w") as f:
        json.dump(out, f, indent=1)
print(json.dumps(out, indent=1)[:3000])
