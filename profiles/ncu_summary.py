"""Summarise an exported `ncu --page raw --csv` file (one or more kernels) into the handful of
metrics the roofline / stall discussion in DESIGN.md uses.  Usage: ncu_summary.py file.raw.csv"""
import csv
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_l1tex_data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for h, u, v in zip(hdr, units, r):
        if h in WANT or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
            try:
                if h.startswith(STALL) and float(v.replace(",", "")) < 0.02:
                    continue
            except ValueError:
                pass
            print(f"{h:90s} {u:16s} {v}")
    print()
