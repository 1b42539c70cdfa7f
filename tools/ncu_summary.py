#!/usr/bin/env python
"""tools/ncu_summary.py <raw.csv from `ncu -i X.ncu-rep --page raw --csv`> <out.csv> "<header comment>"
Transposes the raw page into one row per metric (the ones DESIGN.md cites) and one column per captured launch, plus a sum column
for additive metrics.  Used for profiles/*_ncu_full_*.csv."""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
]
ADDITIVE = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "smsp__inst_executed.sum",
            "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
            "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")].replace("void unnamed>::", "").split("(")[0] for r in data]
    with open(sys.argv[2], "w", newline="") as f:
        f.write(f"# {sys.argv[3]}\n")
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"{n} #{i}" for i, n in enumerate(names)] + ["sum"])
        for m in METRICS:
            if m not in hdr:
                continue
            i = hdr.index(m)
            vals = [r[i] for r in data]
            total = ""
            if m in ADDITIVE and all(num(v) is not None for v in vals):
                total = f"{sum(num(v) for v in vals):.6g}"
            w.writerow([m, units[i]] + vals + [total])


if __name__ == "__main__":
    main()
