#!/usr/bin/env python3
"""Extracts the metrics DESIGN.md / bench.py quote from an .ncu-rep (ncu -i ... --page raw --csv) into a small text file."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# source: %s (ncu --set full --clock-control none), one row per captured launch\n" % rep)
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write("\nkernel: %s\n" % name)
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("  %-84s %s %s\n" % (w, r[i], units[i]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
