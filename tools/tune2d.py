#!/usr/bin/env python3
"""Times jacobi2d pass variants on one GPU (launch-shape sweep).

  python tools/tune2d.py build     # compile all variants (no GPU needed)
  python tools/tune2d.py run       # time them, print a table + JSON lines
"""
import concurrent.futures
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import build as cuda_build  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402

WIDTH = HEIGHT = 16384
ITERATE = 120  # divisible by the time blocks compared below


def variants():
  if os.environ.get('SODA_TUNE_SET') == 'wide':
    # the launch-shape sweep of the first round
    for tb, warps, chunk, stages in itertools.product((1, 2, 4, 8), (4, 8),
                                                      (6, 12), (3, 4)):
      yield tb, {'warps': warps, 'chunk': chunk, 'stages': stages}
    return
  for tb in (6, 7):
    for stages in (2, 3, 4):
      for warps in (2, 4, 8):
        yield tb, {'stages': stages, 'warps': warps}
    yield tb, {'chunk': 3}
    yield tb, {'chunk': 12, 'stages': 2}
    yield tb, {'min_blocks': 3}
    yield tb, {'no_pack': True}
  yield 8, {'cells': 8, 'stages': 2, 'warps': 2}
  yield 8, {}


def stencil():
  with open(os.path.join(ROOT, 'tests', 'src', 'jacobi2d.soda')) as fp:
    return sodac.compile_source(fp.read(), iterate=ITERATE)


def build_all():
  st = stencil()

  def one(v):
    return cuda_build.build_library(st, v[0], v[1])

  with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
    return list(pool.map(one, list(variants())))


def run_all(segments=(0,)):
  import torch
  st = stencil()
  dev = torch.device('cuda', 0)
  d_in = torch.rand((HEIGHT, WIDTH), dtype=torch.float32, device=dev)
  d_out = torch.zeros_like(d_in)
  stream = torch.cuda.current_stream().cuda_stream
  peak = 6535.1
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fp:
      peak = float(json.load(fp)['hbm_gbs'])
  except Exception:  # pylint: disable=broad-except
    pass
  rows = []
  for (tb, options), segment in itertools.product(list(variants()), segments):
    prog = launcher.CudaProgram(cuda_build.build_library(st, tb, options))
    plan = prog.create_plan((WIDTH, HEIGHT),
                            launcher.make_opts(stream=stream, segment=segment))
    pitches = [(WIDTH, 0)]
    run = lambda: plan.run_device([d_in.data_ptr()], pitches,
                                  [d_out.data_ptr()], pitches)
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
    ms = start.elapsed_time(end) / reps
    passes = prog.num_passes
    gcells = WIDTH * HEIGHT * ITERATE / (ms * 1e-3) / 1e9
    gbs = WIDTH * HEIGHT * 8 * passes / (ms * 1e-3) / 1e9
    row = dict(tb=tb, segment=segment, ms_per_pass=ms / passes,
               gcell_per_s=gcells, gbs=gbs, frac=gbs / peak, **options)
    rows.append(row)
    print(json.dumps(row), flush=True)
    plan.close()
  return rows


if __name__ == '__main__':
  if sys.argv[1] == 'build':
    print(len(build_all()), 'variants built')
  else:
    segs = tuple(int(x) for x in sys.argv[2:]) or (0,)
    run_all(segs)
