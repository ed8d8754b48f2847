#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU) into the few metrics the roofline discussion uses.
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "sm__inst_executed.sum", "smsp__inst_executed_op_shared_ld.sum", "sm__cycles_elapsed.max",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
] + ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % x for x in (
    "selected", "not_selected", "wait", "math_pipe_throttle", "mio_throttle", "short_scoreboard", "long_scoreboard", "barrier",
    "no_instruction", "branch_resolving", "dispatch_stall", "lg_throttle")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print(f"# {rep}: {len(data)} captured launch(es)")
    for r in data:
        print(f"\n== {r[ki]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:85s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
