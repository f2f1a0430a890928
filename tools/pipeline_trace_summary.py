#!/usr/bin/env python3
"""Prints the last traced call per chunk layout of a SODA_CUDA_PIPELINE_TRACE=1
run (stderr lines {"pipeline_trace": ...}): when each chunk's upload, compute
and download ended, in ms after the call's first queued operation.

  python tools/pipeline_trace_summary.py trace.err [--chunks]
"""
import json
import sys

last = {}
for line in open(sys.argv[1]):
  if not line.startswith('{"pipeline_trace"'):
    continue
  d = json.loads(line)['pipeline_trace']
  last[(d['chunks'], tuple(d['chunk_slices'][:2]))] = d
for key, d in last.items():
  order = d['compute_order']
  up = d['upload_end_ms']
  print('chunks %d (first slices %s): uploads end %.2f, computes end %.2f, '
        'downloads end %.2f ms' % (key[0], list(key[1]), up[-1],
                                   d['compute_end_ms'][-1], d['download_end_ms'][-1]))
  if '--chunks' in sys.argv:
    for n, k in enumerate(order):
      print('  chunk %2d  %5d slices  upload ends %6.2f  compute ends %6.2f  '
            'download ends %6.2f' % (k, d['chunk_slices'][n], up[k],
                                     d['compute_end_ms'][n], d['download_end_ms'][n]))
