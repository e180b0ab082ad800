"""Compact per-launch summary of an .ncu-rep (read on the CPU box with `ncu -i`): the counters the
roofline discussion in DESIGN.md uses.  usage: python scripts/ncu_summary.py REPORT.ncu-rep > profiles/NAME.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, units = rows[hi], rows[hi + 1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]
for r in rows[hi + 2:]:
    if len(r) < len(hdr):
        continue
    print("==", r[idx["Kernel Name"]][:110])
    for w in WANT:
        if w in idx:
            print(f"   {w:82s} {r[idx[w]]:>18s} {units[idx[w]]}")
