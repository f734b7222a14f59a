"""Turn the ncu captures in gpurun_out/ into the text summaries committed under profiles/ (run in the build container)."""
import csv
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rows = [r for r in csv.reader(open(f"profiles/{tag}_launches_ncu.csv")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(list)
for r in rows[1:]:
    agg[r[ki][:90]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for k, v in agg.items() if "ysb::" in k)
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 80  python bench.py --steps 6 --warmup 3 --no-cpu-baseline",
       "# (cold-cache, serialised launches: compare SHARES, not absolutes)", ""]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    share = f"{100 * sum(v) / tot:5.1f}%" if "ysb::" in k else "  n/a"
    out.append(f"{k:92s} n={len(v):3d} mean_us={sum(v) / len(v) / 1000:8.1f} share_of_ysb_time={share}")
open(f"profiles/{tag}_launch_summary.txt", "w").write("\n".join(out) + "\n")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]
import os
for name, pat in (("filter", "k_filter_planes"), ("nms", "k_select_nms"), ("nms_c3", "k_select_nms"), ("rows", "k_filter_rows")):
    if not os.path.exists(f"gpurun_out/prof_{name}_{tag}.ncu-rep"):
        continue
    txt = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{name}_{tag}.ncu-rep", "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    h = rr[0]
    cfg = {"nms_c3": " --config c3", "rows": " --config c4"}.get(name, "")
    lines = [f"# ncu --set full --clock-control none --import-source on -k regex:{pat} -s 3 -c 2  python bench.py{cfg} --steps 2 --warmup 3 --no-cpu-baseline",
             f"# kernel: {rr[2][h.index('Kernel Name')][:100]}  grid {rr[2][h.index('Grid Size')]} block {rr[2][h.index('Block Size')]}", ""]
    for w in want:
        if w in h:
            i = h.index(w)
            lines.append(f"{w:75s} [{rr[1][i]:>14s}] " + "  ".join(r[i] for r in rr[2:]))
    open(f"profiles/{tag}_{name}_ncu_raw.txt", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:8]))
print(open(f"profiles/{tag}_launch_summary.txt").read()[:900])
