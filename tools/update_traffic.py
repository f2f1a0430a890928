#!/usr/bin/env python3
"""Records the DRAM traffic of one kernel launch (ncu --set full capture) in
profiles/traffic_bytes_per_launch.json under the file name of the program
library the capture ran, which is what bench.py looks up for
`roofline.traffic`: a library rebuilt from changed sources has another hash in
its name and therefore no (stale) entry.

  python tools/update_traffic.py REPORT.ncu-rep LIBRARY.so [SOURCE-NOTE]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, 'profiles', 'traffic_bytes_per_launch.json')
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def dram_bytes(report):
  out = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'],
                       capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units, vals = rows[0], rows[1], rows[2]
  total = 0.0
  for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
    i = hdr.index(name)
    total += float(vals[i].replace(',', '')) * SCALE[units[i]]
  return total


def main():
  report, library = sys.argv[1], os.path.basename(sys.argv[2])
  note = sys.argv[3] if len(sys.argv) > 3 else report
  with open(PATH) as fp:
    table = json.load(fp)
  table[library] = dram_bytes(report)
  table['source_' + library] = note
  with open(PATH, 'w') as fp:
    json.dump(table, fp, indent=1)
    fp.write('\n')
  print(library, table[library])


if __name__ == '__main__':
  main()
