#!/usr/bin/env python3
"""Measured throughput against the planner's model over time blocks.

  python tools/tb_sweep.py build          # compile every variant (no GPU)
  python tools/tb_sweep.py run [program]  # time on cuda:0, one JSON line each

Every variant runs 4 passes of the same time block (iterate = 4 x time block),
device-resident, and is printed next to the model's prediction
(soda_b200/codegen/cuda/model.py) for the same shape.  The output is the
calibration set of tests/test_model.py (profiles/r02_time_block_sweep.jsonl).
"""
import concurrent.futures
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import build as cuda_build  # noqa: E402
from soda_b200.codegen.cuda import launcher, model  # noqa: E402

PASSES = 4


def cases():
  out = []
  for tb in range(1, 9):
    out.append(('jacobi2d', (16384, 16384), tb, {}))
  out.append(('jacobi2d', (16384, 16384), 7, {'cells': 4}))
  out.append(('jacobi2d', (16384, 16384), 8, {'cells': 4}))
  out.append(('jacobi2d', (16384, 16384), 8, {'chunk': 3}))
  for tb in range(1, 9):
    out.append(('seidel2d', (16384, 16384), tb, {}))
  for tb in (4, 5, 6):
    out.append(('seidel2d', (16384, 16384), tb, {'chunk': 3}))
    out.append(('seidel2d', (16384, 16384), tb, {'cells': 4}))
  for tb in (1, 2, 3):
    out.append(('blur', (2000, 16384), tb, {}))
    out.append(('blur', (16000, 16384), tb, {}))
  for name in ('heat3d', 'jacobi3d'):
    for tb in (1, 2, 3, 4):
      out.append((name, (512, 512, 512), tb, {}))
  return out


def stencil(name, tb):
  with open(os.path.join(ROOT, 'tests', 'src', name + '.soda')) as fp:
    return sodac.compile_source(fp.read(), iterate=tb * PASSES)


def build_all():
  def one(case):
    name, _, tb, options = case
    return cuda_build.build_library(stencil(name, tb), tb, options)
  with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
    return list(pool.map(one, cases()))


def run_all(selected):
  import torch
  dev = torch.device('cuda', 0)
  stream = torch.cuda.current_stream().cuda_stream
  for name, extent, tb, options in cases():
    if selected and name not in selected:
      continue
    st = stencil(name, tb)
    prog = launcher.CudaProgram(cuda_build.build_library(st, tb, options))
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
    record = dict(program=name, extent=extent, tb=tb, options=options)
    try:
      for _ in range(3):
        run()
      torch.cuda.synchronize()
      start = torch.cuda.Event(enable_timing=True)
      end = torch.cuda.Event(enable_timing=True)
      reps = 5
      start.record()
      for _ in range(reps):
        run()
      end.record()
      torch.cuda.synchronize()
      ms = start.elapsed_time(end) / reps
      cells = 1
      for e in extent:
        cells *= e
      record.update(ms_per_pass=ms / prog.num_passes,
                    gcells=cells * st.iterate / (ms * 1e-3) / 1e9)
    except launcher.SodaCudaError as e:
      record['error'] = str(e)
    est = model.estimate_pass(st, tb, options, list(extent))
    if est:
      record.update(model_gcells=est['gcells'], model_bound=est['bound'],
                    cells=est['cells'], cy=est['cy'], rows=est['rows'],
                    redundancy=est['redundancy'],
                    instr_per_update=est['instr_per_update'])
    print(json.dumps(record), flush=True)
    plan.close()
    del ins, outs
    torch.cuda.empty_cache()


if __name__ == '__main__':
  if sys.argv[1] == 'build':
    print(len(build_all()), 'libraries built')
  else:
    run_all(set(sys.argv[2:]))
