#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports: the handful of metrics the roofline discussion needs."""
import csv, sys
KEYS = [
 ("gpu__time_duration.sum", "duration"),
 ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
 ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
 ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
 ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/LSU throughput %"),
 ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
 ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
 ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe active %"),
 ("sm__inst_executed_pipe_fp64.sum", "fp64 warp insts"),
 ("smsp__inst_executed.sum", "warp insts"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
 ("launch__registers_per_thread", "registers/thread"), ("launch__block_size", "block"), ("launch__grid_size", "grid"),
 ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
 ("launch__occupancy_limit_registers", "occ limit regs (blocks)"), ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"),
 ("launch__occupancy_limit_warps", "occ limit warps (blocks)"),
 ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 RED sectors"),
 ("lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "L2 atomic unit active %"),
 ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
 ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
 ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
 ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
 ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
 ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
 ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
 ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
 ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
 ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
 ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
]
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if r]
    hi = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    for r in rows[hi + 2:]:
        print(f"## {path}\nkernel: {r[hdr.index('Kernel Name')][:120]}")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {label:28s} {r[i]:>18s} {units[i]}")
