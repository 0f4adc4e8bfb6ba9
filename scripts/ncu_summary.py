"""Condense `ncu -i X.ncu-rep --page raw --csv` into the small per-launch table committed under profiles/.

    ncu -i gpurun_out/step.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/ncu_summary.py /tmp/raw.csv "header comment" > profiles/rN_<kernel>_ncu_full_vK.csv
"""
import csv
import sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size",
    "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__registers_per_thread", "sm__cycles_elapsed.avg.per_second",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if r]
    hdr_i = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr, units, data = rows[hdr_i], rows[hdr_i + 1], rows[hdr_i + 2:]
    col = {n: i for i, n in enumerate(hdr)}
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    print("metric,unit," + ",".join(f"launch{i}" for i in range(len(data))))
    print('"Kernel Name",,' + ",".join('"%s"' % r[col["Kernel Name"]] for r in data))
    for m in KEEP:
        if m in col:
            print(",".join([m, units[col[m]]] + [r[col[m]].replace(",", "") for r in data]))


if __name__ == "__main__":
    main()
