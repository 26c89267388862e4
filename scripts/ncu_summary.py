"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

  python scripts/ncu_summary.py launches <launches.csv> <out.txt>
  python scripts/ncu_summary.py full <prof.ncu-rep> <out.txt>
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu.sum", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        try:
            val = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[ui], 1e-6)
        name = r[ki].split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += val * scale
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare shares)\n")
        f.write("# source: %s\n" % path)
        f.write("%-72s %6s %12s %10s %7s\n" % ("kernel", "n", "total_ms", "avg_us", "share"))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %6d %12.4f %10.2f %7.4f\n" % (k, v[0], v[1], 1e3 * v[1] / v[0], v[1] / tot))
        f.write("%-72s %6s %12.4f\n" % ("TOTAL", "", tot))
    print(open(out).read())


def full(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on; source: %s\n" % path)
        ki = h.index("Kernel Name")
        for r in rows[2:]:
            f.write("\n## launch id %s: %s\n" % (r[0], r[ki][:100]))
            for m in METRICS:
                if m in h:
                    i = h.index(m)
                    f.write("%-85s %18s %s\n" % (m, r[i], rows[1][i]))
    print(open(out).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
