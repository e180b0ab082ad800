"""Issue-bound accounting of DESIGN.md section 4 from an ncu summary (scripts/ncu_summary.py output):
per kernel (first launch of each name): warp instructions, fp64 warp instructions (fp64 pipe busy share x active
cycles / 2), the bound 2*N64 + Nx in cycles per scheduler and in microseconds at the measured clock, and DRAM bytes."""
import re, sys
txt = open(sys.argv[1]).read()
SCHED = 148 * 4
seen = set()
for blk in txt.split("== ")[1:]:
    name = blk.split("(")[0].split("::")[-1].strip()
    full = blk.splitlines()[0][:70]
    if full in seen:
        continue
    seen.add(full)
    def g(key):
        m = re.search(re.escape(key) + r"\s+([0-9.,]+)", blk)
        return float(m.group(1).replace(",", "")) if m else float("nan")
    t_us = g("gpu__time_duration.sum")
    inst = g("smsp__inst_executed.sum")
    act = g("smsp__cycles_active.avg")
    el = g("sm__cycles_elapsed.max")
    f64 = g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") / 100.0
    n64 = f64 * act / 2.0
    nx = inst / SCHED - n64
    bound = 2 * n64 + nx
    ghz = el / t_us / 1e3
    print(f"{full}\n   t={t_us:.1f} us  inst={inst/1e6:.2f} M  fp64={n64*SCHED/1e6:.2f} M  bound={bound/1e3:.1f} kcyc = {bound/ghz/1e3:.1f} us"
          f"  active={act/1e3:.1f} kcyc  measured/bound={bound/act:.2f}  fp64 pipe={100*f64:.0f} %  regs={g('launch__registers_per_thread'):.0f}"
          f"  warps_active={g('sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} %"
          f"  dram={(g('dram__bytes_read.sum')+g('dram__bytes_write.sum')):.1f} MB  clock={ghz:.2f} GHz")
