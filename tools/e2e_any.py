#!/usr/bin/env python3
"""End-to-end (pinned host arrays) time of any program through
soda_cuda_plan_run_host for several chunk counts of the host pipeline.

  python tools/e2e_any.py heat3d 512,512,512 --iterate 32 --chunks 1 2 4 8 0
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import build as cuda_build  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('program')
ap.add_argument('extent')
ap.add_argument('--iterate', type=int, default=None)
ap.add_argument('--tb', type=int, default=None)
ap.add_argument('--chunks', type=int, nargs='+', default=[1, 0])
ap.add_argument('--build-only', action='store_true')
ap.add_argument('--one-shot', action='store_true',
                help='also time the soda_cuda_<app> entry (a new plan, with its '
                'device arrays, for every call)')
args = ap.parse_args()
extent = tuple(int(x) for x in args.extent.split(','))
overrides = {'iterate': args.iterate} if args.iterate else {}
with open(os.path.join(ROOT, 'tests', 'src', args.program + '.soda')) as fp:
  st = sodac.compile_source(fp.read(), **overrides)
lib = cuda_build.build_library(st, args.tb, {})
if args.build_only:
  print(lib)
  sys.exit(0)
prog = launcher.CudaProgram(lib)
shape = extent[::-1]
rng = np.random.default_rng(0)
buffers, inputs, outputs = [], {}, {}
for name, dt in zip(prog.input_names, prog.input_dtypes):
  buf = launcher.HostBuffer(prog, shape, dt, 0)
  if np.dtype(dt).kind == 'f':
    buf.array[...] = rng.random(shape, dtype=np.float32).astype(dt)
  else:
    buf.array[...] = rng.integers(0, 1000, shape).astype(dt)
  buffers.append(buf)
  inputs[name] = buf.array
for name, dt in zip(prog.output_names, prog.output_dtypes):
  buf = launcher.HostBuffer(prog, shape, dt, 0)
  buffers.append(buf)
  outputs[name] = buf.array
nbytes = sum(a.nbytes for a in inputs.values()) + sum(a.nbytes for a in outputs.values())
for chunks in args.chunks:
  plan = prog.create_plan(extent, launcher.make_opts(host_chunks=chunks))
  plan.run_host(inputs, outputs)
  torch.cuda.synchronize()
  times = []
  for _ in range(5):
    t0 = time.perf_counter()
    plan.run_host(inputs, outputs)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
  best, mean = min(times), sum(times) / len(times)
  print(json.dumps(dict(program=args.program, extent=extent, iterate=st.iterate,
                        passes=prog.num_passes, chunks=chunks,
                        ms_best=best * 1e3, ms_mean=mean * 1e3,
                        gbs_mean=nbytes / mean / 1e9)), flush=True)
  plan.close()
if args.one_shot:
  prog.run_host(inputs, outputs)
  torch.cuda.synchronize()
  times = []
  for _ in range(5):
    t0 = time.perf_counter()
    prog.run_host(inputs, outputs)
    times.append(time.perf_counter() - t0)
  best, mean = min(times), sum(times) / len(times)
  print(json.dumps(dict(program=args.program, extent=extent, iterate=st.iterate,
                        passes=prog.num_passes, chunks='soda_cuda_<app> (one-shot)',
                        ms_best=best * 1e3, ms_mean=mean * 1e3,
                        gbs_mean=nbytes / mean / 1e9)), flush=True)
