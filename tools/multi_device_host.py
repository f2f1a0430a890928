#!/usr/bin/env python3
"""One process, several devices: soda_cuda_multi_run_host (what `sodac
--cuda-gpus N` puts behind the program-named entry point) on the bench grid,
jacobi2d fp32 16384 x (16384 N) iterate 64, host arrays in pinned memory, for
N = 1, 2, 4, ... visible devices.  Prints one JSON line per N with the
end-to-end rate and a dependency-cone parity check on windows that straddle
the device boundaries."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import cone  # noqa: E402
from soda_b200.codegen.cuda import launcher, multi_gpu  # noqa: E402


def main():
  st, prog = bench.config_program(bench.HEADLINE)
  width, height = bench.WIDTH, bench.HEIGHT
  n = 1
  while n <= torch.cuda.device_count():
    extent = (width, height * n)
    pinned_in = launcher.HostBuffer(prog, extent[::-1], np.float32, 0)
    pinned_out = launcher.HostBuffer(prog, extent[::-1], np.float32, 0)
    x = torch.from_numpy(pinned_in.array)
    block = 4096
    for r0 in range(0, extent[1], block):
      x[r0:r0 + block].copy_(bench.synthetic(((0, width), (r0, r0 + block)), 0,
                                             torch.float32,
                                             torch.device('cuda', 0)))
    inputs, outputs = {'t1': pinned_in.array}, {'t0': pinned_out.array}
    multi_gpu.run_host_multi(prog, inputs, outputs, num_devices=n)  # warm-up
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
      multi_gpu.run_host_multi(prog, inputs, outputs, num_devices=n)
    seconds = (time.perf_counter() - t0) / reps
    seams = [(width // 2, height * k - 32) for k in range(1, n)]
    report = cone.check_host_arrays(st, inputs, outputs, count=4, seed=n,
                                    required=seams)
    print(json.dumps({
        'devices': n, 'extent': extent, 'iterate': st.iterate,
        'ms': seconds * 1e3,
        'gcell_per_s': extent[0] * extent[1] * st.iterate / seconds / 1e9,
        'host_gbs': 2 * extent[0] * extent[1] * 4 / seconds / 1e9,
        'parity': {k: report[k] for k in ('windows', 'bit_exact')},
    }), flush=True)
    pinned_in.close()
    pinned_out.close()
    n *= 2


if __name__ == '__main__':
  main()
