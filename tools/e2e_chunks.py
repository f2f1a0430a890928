#!/usr/bin/env python3
"""e2e (host arrays) time of the bench workload for several pipeline chunk
counts, plus device-resident time for a few time blocks."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from soda_b200 import sodac
from soda_b200.codegen import cuda as cb
from soda_b200.codegen.cuda import launcher

W = H = 16384
st = sodac.compile_source(open(os.path.join(ROOT, 'tests/src/jacobi2d.soda')).read(), iterate=64)
for tb in (4, 8):
  prog = cb.compile_stencil(st, time_block=tb)
  h_in = torch.rand((H, W), dtype=torch.float32).pin_memory()
  h_out = torch.zeros((H, W), dtype=torch.float32).pin_memory()
  for chunks in (1, 4, 8, 16, 32):
    plan = prog.create_plan((W, H), launcher.make_opts(host_chunks=chunks))
    plan.run_host({'t1': h_in.numpy()}, {'t0': h_out.numpy()})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
      plan.run_host({'t1': h_in.numpy()}, {'t0': h_out.numpy()})
    dt = (time.perf_counter() - t0) / 3
    print(json.dumps(dict(tb=tb, chunks=chunks, ms=dt * 1e3, gcell=W * H * 64 / dt / 1e9)), flush=True)
    plan.close()
  d_in = torch.rand((H, W), dtype=torch.float32, device='cuda')
  d_out = torch.zeros_like(d_in)
  plan = prog.create_plan((W, H))
  for _ in range(2):
    plan.run_device([d_in.data_ptr()], [(W, 0)], [d_out.data_ptr()], [(W, 0)])
  torch.cuda.synchronize()
  s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s.record()
  for _ in range(5):
    plan.run_device([d_in.data_ptr()], [(W, 0)], [d_out.data_ptr()], [(W, 0)])
  e.record(); torch.cuda.synchronize()
  ms = s.elapsed_time(e) / 5
  print(json.dumps(dict(tb=tb, device_ms=ms, ms_per_pass=ms / prog.num_passes, gcell=W * H * 64 / ms / 1e6)), flush=True)
  plan.close()
