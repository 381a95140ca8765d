#!/usr/bin/env python
"""Boil an .ncu-rep (ncu --set full) down to the metrics profiles/*.csv carry.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [launch_index] > profiles/rNN_ncu_<kernel>_metrics.csv"""
import csv, subprocess, sys
WANT = """gpu__time_duration.sum sm__cycles_elapsed.avg.per_second sm__cycles_active.avg launch__grid_size launch__block_size
launch__registers_per_thread launch__shared_mem_per_block_dynamic launch__waves_per_multiprocessor launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active sm__throughput.avg.pct_of_peak_sustained_elapsed
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__t_bytes.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum smsp__inst_executed.sum
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_membar_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio""".split()
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
r = data[idx]
print("metric,unit,value")
print(f'kernel,,"{r[hdr.index("Kernel Name")]}"')
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w},{units[i]},{r[i]}")
