#!/usr/bin/env python3
"""Runs a few programs through the CPU emulation built with AddressSanitizer
and UBSan: out-of-bounds shared-memory / global accesses of the kernel
templates and the runtime show up as sanitizer reports (compute-sanitizer is
not available on the GPU pool).

  LD_PRELOAD=$(gcc -print-file-name=libasan.so) python tests/emu/run_sanitized.py

(ThreadSanitizer is of no use here: halo cells of a tile are computed from rows
outside the tile, i.e. from whatever lies next to the plane in shared memory,
on purpose - their values are never stored - and every such read races with
some export.)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402
from tests import common  # noqa: E402
from tests.emu import build_emu  # noqa: E402

CASES = [
    ('tests/src/jacobi3d.soda', dict(iterate=3), 2, {'rows': 16, 'cy': 4}, (150, 25, 11)),
    ('tests/src/heat3d.soda', dict(iterate=3), 3, {'rows': 16, 'cy': 2}, (40, 30, 9)),
    ('tests/src/heat3d.soda', dict(iterate=2), 1, {}, (130, 9, 7)),
    ('tests/src/denoise3d.soda', {}, None, {}, (140, 19, 7)),
    ('tests/src/jacobi2d.soda', dict(iterate=7), 6, {}, (300, 40)),
    ('tests/src/jacobi2d.soda', dict(iterate=2), None, {}, (3, 3)),
    ('tests/src/blur.soda', dict(iterate=2), 2, {}, (700, 33)),
    ('tests/src/contrast.soda', {}, None, {}, (200, 40)),
    ('tests/src/xcorr.soda', {}, None, {}, (300, 30)),
    ('tests/src_extra/narrow2d.soda', {}, 2, {}, (90, 21)),
    ('tests/src_extra/conv2d_param.soda', {}, None, {}, (70, 20)),
]


def main():
  mode = True
  for path, overrides, tb, options, extent in CASES:
    with open(os.path.join(ROOT, path)) as fp:
      st = sodac.compile_source(fp.read(), **overrides)
    lib = build_emu.build_emu_library(st, time_block=tb, options=options,
                                      sanitize=mode)
    prog = launcher.CudaProgram(lib)
    rng = np.random.default_rng(1)
    inputs = {}
    for name, dtype in zip(prog.input_names, prog.input_dtypes):
      if dtype.kind == 'f':
        inputs[name] = rng.random(extent[::-1]).astype(dtype)
      else:
        inputs[name] = rng.integers(0, 200, extent[::-1]).astype(dtype)
    params = common.make_params(st) if st.param_stmts else None
    for chunks in (0, 3):
      prog.run_host(inputs, params=params,
                    opts=launcher.make_opts(host_chunks=chunks))
    print('ok', os.path.basename(path), tb, options, extent, flush=True)


if __name__ == '__main__':
  main()
