#!/usr/bin/env python3
"""Device-resident timing of the other BASELINE configs (C1, C3, C4) and of
launch-shape variants.  `build` compiles (no GPU), `run` times on cuda:0 and
prints one JSON line per case with Gcell-updates/s and the HBM roofline
fraction (algorithmic bytes per pass / time per pass / measured copy peak)."""
import concurrent.futures
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import build as cuda_build  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402


def cases():
  """(program, Stencil overrides, grid, time block, options); empty options =
  the planner's defaults."""
  out = []
  out.append(('blur', {'iterate': 2}, (2000, 16384), 2, {}))
  out.append(('blur', {'iterate': 2}, (2000, 16384), 1, {}))
  out.append(('blur', {'iterate': 2}, (16000, 16384), 2, {}))
  for name in ('heat3d', 'jacobi3d'):
    for tb in (1, 2, 3):
      out.append((name, {'iterate': 32}, (512, 512, 512), tb, {}))
  out.append(('denoise3d', {}, (512, 512, 512), 1, {}))
  out.append(('denoise2d', {}, (8192, 8192), 1, {}))
  out.append(('seidel2d', {'iterate': 16}, (16384, 16384), 4, {}))
  out.append(('seidel2d', {'iterate': 16}, (16384, 16384), 4, {'no_pack': True}))
  out.append(('sobel2d', {}, (16384, 16384), 1, {}))
  out.append(('contrast', {}, (16384, 16384), 1, {}))
  out.append(('contrast', {}, (16384, 16384), 1, {'no_pack': True}))
  out.append(('erosion', {}, (16384, 16384), 1, {}))
  out.append(('erosion', {}, (16384, 16384), 1, {'row_unroll': 6}))
  out.append(('xcorr', {}, (16384, 16384), 1, {}))
  out.append(('xcorr', {}, (16384, 16384), 1, {'row_unroll': 6}))
  for tb in (5, 6, 8):
    out.append(('jacobi2d', {'iterate': 120}, (16384, 16384), tb, {}))
  return out


def stencil(name, overrides):
  with open(os.path.join(ROOT, 'tests', 'src', name + '.soda')) as fp:
    return sodac.compile_source(fp.read(), **overrides)


def build_all():
  def one(case):
    name, overrides, _, tb, options = case
    return cuda_build.build_library(stencil(name, overrides), tb, options)
  with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
    return list(pool.map(one, cases()))


def run_all(selected=None):
  import torch
  peak = 6535.1
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fp:
      peak = float(json.load(fp)['hbm_gbs'])
  except Exception:  # pylint: disable=broad-except
    pass
  dev = torch.device('cuda', 0)
  stream = torch.cuda.current_stream().cuda_stream
  for case in cases():
    name, overrides, extent, tb, options = case
    if selected and name not in selected:
      continue
    st = stencil(name, overrides)
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
    try:
      for _ in range(2):
        run()
      torch.cuda.synchronize()
      start = torch.cuda.Event(enable_timing=True)
      end = torch.cuda.Event(enable_timing=True)
      reps = 3
      start.record()
      for _ in range(reps):
        run()
      end.record()
      torch.cuda.synchronize()
    except launcher.SodaCudaError as e:
      print(json.dumps(dict(program=name, tb=tb, options=options,
                            error=str(e))), flush=True)
      continue
    ms = start.elapsed_time(end) / reps
    cells = 1
    for e in extent:
      cells *= e
    passes = prog.num_passes
    gbs = cells * prog.bytes_per_cell_per_pass * passes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(program=name, extent=extent, iterate=st.iterate,
                          tb=tb, options=options, passes=passes,
                          ms_per_pass=ms / passes,
                          gcell_per_s=cells * st.iterate / (ms * 1e-3) / 1e9,
                          gbs=gbs, frac=gbs / peak)), flush=True)
    plan.close()
    del ins, outs
    torch.cuda.empty_cache()


if __name__ == '__main__':
  if sys.argv[1] == 'build':
    print(len(build_all()), 'libraries built')
  else:
    run_all(set(sys.argv[2:]))
