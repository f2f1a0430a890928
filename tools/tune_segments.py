#!/usr/bin/env python3
"""Sweeps the segment length (slices per CTA along the streamed dimension) for
a few kernels, to calibrate choose_segment() in soda_runtime.cuh.
`build` compiles, `run` times on cuda:0 (segment 0 = the runtime's choice)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import build as cuda_build  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402

CASES = [
    ('jacobi2d', 120, (16384, 16384), 5, {},
     (0, 96, 128, 171, 256, 342, 512, 683, 1024, 1366, 2048, 4096)),
    ('jacobi2d', 120, (16384, 16384), 6, {},
     (0, 96, 128, 171, 256, 342, 512, 683, 1024, 1366, 2048, 4096)),
    ('jacobi2d', 120, (16384, 16384), 8, {},
     (0, 128, 256, 512, 1024, 2048)),
    ('jacobi3d', 12, (512, 512, 512), 1, {},
     (0, 16, 24, 32, 43, 64, 86, 128, 171, 256, 512)),
    ('jacobi3d', 12, (512, 512, 512), 1, {'pack': True},
     (0, 32, 64, 128, 256)),
    ('heat3d', 12, (512, 512, 512), 1, {},
     (0, 32, 64, 128, 256)),
    ('jacobi3d', 12, (512, 512, 512), 2, {},
     (0, 16, 24, 32, 43, 64, 86, 128, 171, 256, 512)),
    ('heat3d', 12, (512, 512, 512), 2, {},
     (0, 32, 64, 86, 128, 256)),
]


def stencil(name, iterate):
  with open(os.path.join(ROOT, 'tests', 'src', name + '.soda')) as fp:
    return sodac.compile_source(fp.read(), iterate=iterate)


def main():
  if sys.argv[1] == 'build':
    for name, iterate, _, tb, options, _ in CASES:
      print(cuda_build.build_library(stencil(name, iterate), tb, options))
    return
  import torch
  dev = torch.device('cuda', 0)
  stream = torch.cuda.current_stream().cuda_stream
  for name, iterate, extent, tb, options, segments in CASES:
    st = stencil(name, iterate)
    prog = launcher.CudaProgram(cuda_build.build_library(st, tb, options))
    shape = tuple(extent[::-1])
    d_in = torch.rand(shape, dtype=torch.float32, device=dev)
    d_out = torch.zeros_like(d_in)
    plane = extent[0] * extent[1] if len(extent) == 3 else 0
    pitches = [(extent[0], plane)]
    cells = 1
    for e in extent:
      cells *= e
    for segment in segments:
      plan = prog.create_plan(extent, launcher.make_opts(stream=stream,
                                                         segment=segment))
      run = lambda: plan.run_device([d_in.data_ptr()], pitches,
                                    [d_out.data_ptr()], pitches)
      for _ in range(2):
        run()
      torch.cuda.synchronize()
      start = torch.cuda.Event(enable_timing=True)
      end = torch.cuda.Event(enable_timing=True)
      start.record()
      for _ in range(3):
        run()
      end.record()
      torch.cuda.synchronize()
      ms = start.elapsed_time(end) / 3
      print(json.dumps(dict(program=name, tb=tb, options=options,
                            segment=segment,
                            ms_per_pass=ms / prog.num_passes,
                            gcell_per_s=cells * iterate / (ms * 1e-3) / 1e9)),
            flush=True)
      plan.close()
    del d_in, d_out
    torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
