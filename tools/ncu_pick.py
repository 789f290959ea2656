"""Pick the metrics the profiles/ summaries quote out of an `ncu --page raw --csv`
export (one kernel launch per row)."""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        u = dict(zip(names, units))
        out.append({"kernel": d.get("Kernel Name", "")[:120],
                    "metrics": {k: d[k] + (" " + u[k] if u.get(k) else "") for k in KEYS if k in d}})
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
