#!/usr/bin/env python3
"""Times 3-D pass variants (tile rows, rows per thread, time block) of
jacobi3d / heat3d / denoise3d on one GPU.

  python tools/tune3d.py build     # compile all variants (no GPU needed)
  python tools/tune3d.py run [program ...]
"""
import concurrent.futures
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import build as cuda_build  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402

EXTENT = (512, 512, 512)
ITERATE = 12


def variants():
  out = []
  if os.environ.get('SODA_TUNE_SET') == 'final':
    for name in ('jacobi3d', 'heat3d'):
      for rows, cy in ((16, 4), (24, 4), (32, 4), (16, 2), (32, 2), (12, 4)):
        for mb in (2, 3, 4):
          out.append((name, {'iterate': ITERATE}, 2,
                      {'rows': rows, 'cy': cy, 'min_blocks': mb}))
      out.append((name, {'iterate': ITERATE}, 2, {'pack': True}))
      out.append((name, {'iterate': ITERATE}, 2, {}))
      out.append((name, {'iterate': ITERATE}, 1, {}))
    for rows, cy in ((20, 1), (24, 1), (28, 1), (32, 1)):
      for mb in (1,):
        out.append(('denoise3d', {}, 1,
                    {'rows': rows, 'cy': cy, 'min_blocks': mb}))
    return out
  if os.environ.get('SODA_TUNE_SET') == 'tb2':
    for name in ('jacobi3d', 'heat3d'):
      for rows, cy in ((16, 2), (32, 2), (32, 4), (16, 4), (24, 4), (48, 4)):
        for mb in (2, 3, 4):
          for extra in ({}, {'no_pack': True}):
            opts = {'rows': rows, 'cy': cy, 'min_blocks': mb}
            opts.update(extra)
            out.append((name, {'iterate': ITERATE}, 2, opts))
      for rows, cy in ((8, 2), (8, 4), (12, 4), (16, 4)):
        for mb in (4, 6):
          out.append((name, {'iterate': ITERATE}, 1,
                      {'rows': rows, 'cy': cy, 'min_blocks': mb}))
    return out
  for name in ('jacobi3d', 'heat3d'):
    for tb, shapes in (
        (1, [(8, 1), (8, 2), (16, 2), (16, 4), (32, 4), (32, 2), (64, 4)]),
        (2, [(32, 1), (16, 2), (32, 2), (32, 4), (64, 4), (64, 2)]),
        (3, [(32, 2), (32, 4), (64, 4)]),
        (4, [(32, 4), (64, 4)]),
    ):
      for rows, cy in shapes:
        out.append((name, {'iterate': ITERATE}, tb, {'rows': rows, 'cy': cy}))
        if tb >= 2 and cy == 4:
          out.append((name, {'iterate': ITERATE}, tb,
                      {'rows': rows, 'cy': cy, 'min_blocks': 2}))
    out.append((name, {'iterate': ITERATE}, 2,
                {'rows': 32, 'cy': 4, 'no_pack': True}))
  for rows, cy in ((16, 1), (16, 2), (32, 2), (32, 4), (16, 4), (64, 4)):
    out.append(('denoise3d', {}, 1, {'rows': rows, 'cy': cy}))
  return out


def stencil(name, overrides):
  with open(os.path.join(ROOT, 'tests', 'src', name + '.soda')) as fp:
    return sodac.compile_source(fp.read(), **overrides)


def build_all():
  def one(v):
    name, overrides, tb, options = v
    try:
      return cuda_build.build_library(stencil(name, overrides), tb, options)
    except Exception as e:  # pylint: disable=broad-except
      return 'FAILED %s %s %s: %s' % (name, tb, options, str(e)[:300])
  with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
    return list(pool.map(one, variants()))


def run_all(selected):
  import torch
  peak = 6535.1
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fp:
      peak = float(json.load(fp)['hbm_gbs'])
  except Exception:  # pylint: disable=broad-except
    pass
  dev = torch.device('cuda', 0)
  stream = torch.cuda.current_stream().cuda_stream
  shape = EXTENT[::-1]
  cells = EXTENT[0] * EXTENT[1] * EXTENT[2]
  bufs = [torch.rand(shape, dtype=torch.float32, device=dev) for _ in range(3)]
  for name, overrides, tb, options in variants():
    if selected and name not in selected:
      continue
    st = stencil(name, overrides)
    try:
      prog = launcher.CudaProgram(cuda_build.build_library(st, tb, options))
      plan = prog.create_plan(EXTENT, launcher.make_opts(stream=stream))
      ins = bufs[:len(prog.input_dtypes)]
      outs = bufs[len(prog.input_dtypes):][:1]
      pitches = [(EXTENT[0], EXTENT[0] * EXTENT[1])]
      run = lambda: plan.run_device([t.data_ptr() for t in ins],
                                    pitches * len(ins),
                                    [t.data_ptr() for t in outs], pitches)
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
    except Exception as e:  # pylint: disable=broad-except
      print(json.dumps(dict(program=name, tb=tb, options=options,
                            error=str(e)[:200])), flush=True)
      continue
    ms = start.elapsed_time(end) / reps
    passes = prog.num_passes
    gbs = cells * prog.bytes_per_cell_per_pass * passes / (ms * 1e-3) / 1e9
    print(json.dumps(dict(program=name, tb=tb, options=options, passes=passes,
                          ms_per_pass=ms / passes,
                          gcell_per_s=cells * st.iterate / (ms * 1e-3) / 1e9,
                          frac=gbs / peak)), flush=True)
    plan.close()


if __name__ == '__main__':
  if sys.argv[1] == 'build':
    libs = build_all()
    bad = [l for l in libs if l.startswith('FAILED')]
    print(len(libs) - len(bad), 'variants built;', len(bad), 'failed')
    for b in bad:
      print(b)
  else:
    run_all(set(sys.argv[2:]))
