#!/usr/bin/env python3
"""End-to-end (pinned host arrays) time of the bench workload through
soda_cuda_plan_run_host for several chunk layouts of the host pipeline:
n > 0 equal chunks, n < 0 ramped (quarter / half chunks at both ends), 0 auto.

  python tools/e2e_ab.py [chunks ...]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402

W, H = bench.WIDTH, bench.HEIGHT
st, prog = bench.config_program(bench.HEADLINE)
layouts = [int(x) for x in sys.argv[1:]] or [16, -16, 0, -20, -24, 16, 0]
pinned_in = launcher.HostBuffer(prog, (H, W), np.float32, 0)
pinned_out = launcher.HostBuffer(prog, (H, W), np.float32, 0)
np_in, np_out = pinned_in.array, pinned_out.array
np_in[:] = np.random.default_rng(0).random((H, W), dtype=np.float32)
for chunks in layouts:
  plan = prog.create_plan((W, H), launcher.make_opts(host_chunks=chunks))
  plan.run_host({'t1': np_in}, {'t0': np_out})
  torch.cuda.synchronize()
  times = []
  for _ in range(5):
    t0 = time.perf_counter()
    plan.run_host({'t1': np_in}, {'t0': np_out})
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
  best, mean = min(times), sum(times) / len(times)
  print(json.dumps(dict(chunks=chunks, ms_best=best * 1e3, ms_mean=mean * 1e3,
                        gcell_mean=W * H * bench.ITERATE / mean / 1e9,
                        gbs_mean=2 * W * H * 4 / mean / 1e9)), flush=True)
  plan.close()
