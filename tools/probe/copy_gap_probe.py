#!/usr/bin/env python3
"""What does cutting a pinned host<->device transfer into pieces cost?

1 GiB up and 1 GiB down at the same time (two streams, like the copy-in and
copy-out streams of the host pipeline), cut into pieces of several sizes, with
and without an event record between the pieces, and with the pieces of one
direction alternating over two streams.  Prints one JSON line per layout.

  python tools/probe/copy_gap_probe.py
"""
import json
import time

import torch

GIB = 1 << 30
dev = torch.device('cuda', 0)
h_in = torch.empty(GIB, dtype=torch.uint8).pin_memory()
h_out = torch.empty(GIB, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in = torch.empty(GIB, dtype=torch.uint8, device=dev)
d_out = torch.ones(GIB, dtype=torch.uint8, device=dev)
up = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
down = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]


def run(piece_mib, events, two_streams, directions='both', lag=1):
  piece = piece_mib << 20
  n = GIB // piece
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  marks = []
  for k in range(n + lag):
    lo = k * piece
    if k < n and directions in ('both', 'up'):
      s = up[k % 2 if two_streams else 0]
      with torch.cuda.stream(s):
        d_in[lo:lo + piece].copy_(h_in[lo:lo + piece], non_blocking=True)
        if events:
          e = torch.cuda.Event()
          e.record(s)
          marks.append(e)
    j = k - lag  # downloads trail the uploads by `lag` pieces, as in the pipeline
    if j >= 0 and directions in ('both', 'down'):
      s = down[j % 2 if two_streams else 0]
      lo = j * piece
      with torch.cuda.stream(s):
        if events and directions == 'both':
          s.wait_event(marks[j])
        h_out[lo:lo + piece].copy_(d_out[lo:lo + piece], non_blocking=True)
        if events:
          e = torch.cuda.Event()
          e.record(s)
  torch.cuda.synchronize()
  return time.perf_counter() - t0


for directions in ('up', 'down', 'both'):
  for piece_mib in (1024, 64, 32, 16, 8):
    for events in (False, True):
      for two in (False, True):
        if piece_mib == 1024 and (events or two):
          continue
        run(piece_mib, events, two, directions)
        times = [run(piece_mib, events, two, directions) for _ in range(4)]
        best = min(times)
        nbytes = GIB * (2 if directions == 'both' else 1)
        print(json.dumps(dict(directions=directions, piece_mib=piece_mib,
                              events=events, two_streams=two,
                              ms_best=best * 1e3,
                              ms_mean=sum(times) / len(times) * 1e3,
                              gbs=nbytes / best / 1e9)), flush=True)
