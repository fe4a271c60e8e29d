#!/usr/bin/env python3
"""Selected metrics of .ncu-rep files as a small CSV (for profiles/).

  python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/x.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum',
    'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
    'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__waves_per_multiprocessor',
    'smsp__average_warp_latency_issue_stalled_barrier.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
]


def main():
  out = csv.writer(sys.stdout)
  out.writerow(['report', 'kernel', 'metric', 'unit', 'value'])
  for path in sys.argv[1:]:
    text = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'],
                          stdout=subprocess.PIPE, text=True,
                          check=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units = rows[0], rows[1]
    for row in rows[2:]:
      record = dict(zip(header, row))
      kernel = record.get('Kernel Name', '')[:60]
      for name in KEEP:
        if name in record:
          out.writerow([path.split('/')[-1], kernel, name,
                        units[header.index(name)], record[name]])


if __name__ == '__main__':
  main()
