#!/usr/bin/env python3
"""Prints the metrics we track from an .ncu-rep (ncu --set full) as text.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...]
"""
import csv
import subprocess
import sys

WANT = [
    'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg.per_second',
    'smsp__cycles_active.avg',
]
STALL = 'smsp__average_warps_issue_stalled_'


def main():
  for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
      print('==', path, '::', vals[hdr.index('Kernel Name')][:90])
      for name in WANT:
        if name in hdr:
          i = hdr.index(name)
          print('%-72s %-14s %s' % (name, units[i], vals[i]))
      for i, name in enumerate(hdr):
        if name.startswith(STALL) and name.endswith('_per_issue_active.ratio'):
          try:
            v = float(vals[i])
          except ValueError:
            continue
          if v >= 0.05:
            print('%-72s %-14s %.3f' % ('stall ' + name[len(STALL):-23], '', v))


if __name__ == '__main__':
  main()
