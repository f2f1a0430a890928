#!/usr/bin/env python3
"""Runs one program variant on cuda:0 with device-resident synthetic inputs and
prints one JSON line (used under ncu and for quick A/B timing).

  python tools/run_one.py jacobi3d 512,512,512 --iterate 32 --tb 2 \
      --options '{"rows": 32}' --reps 3
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import build as cuda_build  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('program')
  ap.add_argument('extent')
  ap.add_argument('--iterate', type=int, default=None)
  ap.add_argument('--tb', type=int, default=None)
  ap.add_argument('--options', default='{}')
  ap.add_argument('--reps', type=int, default=3)
  ap.add_argument('--warmup', type=int, default=2)
  ap.add_argument('--cr', default=None)
  ap.add_argument('--build-only', action='store_true')
  args = ap.parse_args()
  extent = tuple(int(x) for x in args.extent.split(','))
  overrides = {}
  if args.iterate:
    overrides['iterate'] = args.iterate
  if args.cr:
    overrides['computation_reuse'] = args.cr
  path = os.path.join(ROOT, 'tests', 'src', args.program + '.soda')
  if not os.path.exists(path):  # programs of our own (params, widths, half)
    path = os.path.join(ROOT, 'tests', 'src_extra', args.program + '.soda')
  with open(path) as fp:
    st = sodac.compile_source(fp.read(), **overrides)
  options = json.loads(args.options)
  lib = cuda_build.build_library(st, args.tb, options)
  if args.build_only:
    print(lib)
    return
  import torch
  prog = launcher.CudaProgram(lib)
  dev = torch.device('cuda', 0)
  stream = torch.cuda.current_stream().cuda_stream
  shape = tuple(extent[::-1])
  ins, outs = [], []
  for dt in prog.input_dtypes:
    tdt = getattr(torch, str(dt))
    if dt.kind == 'f':
      ins.append(torch.rand(shape, dtype=tdt, device=dev))
    else:
      ins.append(torch.randint(0, 1000, shape, device=dev).to(tdt))
  for dt in prog.output_dtypes:
    outs.append(torch.zeros(shape, dtype=getattr(torch, str(dt)), device=dev))
  plane = extent[0] * extent[1] if len(extent) == 3 else 0
  pitches = [(extent[0], plane)]
  plan = prog.create_plan(extent, launcher.make_opts(stream=stream))
  run = lambda: plan.run_device([t.data_ptr() for t in ins],
                                pitches * len(ins),
                                [t.data_ptr() for t in outs],
                                pitches * len(outs))
  for _ in range(args.warmup):
    run()
  torch.cuda.synchronize()
  start = torch.cuda.Event(enable_timing=True)
  end = torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(args.reps):
    run()
  end.record()
  torch.cuda.synchronize()
  ms = start.elapsed_time(end) / args.reps
  cells = 1
  for e in extent:
    cells *= e
  passes = prog.num_passes
  peak = 6535.1
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fp:
      peak = float(json.load(fp)['hbm_gbs'])
  except Exception:  # pylint: disable=broad-except
    pass
  gbs = cells * prog.bytes_per_cell_per_pass * passes / (ms * 1e-3) / 1e9
  print(json.dumps(dict(program=args.program, extent=extent,
                        iterate=st.iterate, tb=args.tb, options=options,
                        cr=args.cr, passes=passes, ms_per_pass=ms / passes,
                        gcell_per_s=cells * st.iterate / (ms * 1e-3) / 1e9,
                        gbs=gbs, frac=gbs / peak)), flush=True)
  plan.close()


if __name__ == '__main__':
  main()
