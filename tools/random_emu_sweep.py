#!/usr/bin/env python3
"""The differential sweep of tools/random_sweep.py under CPU emulation (no GPU):
   python tools/random_emu_sweep.py LO HI [--hard]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import emit_cpp, golden  # noqa: E402
from soda_b200 import sodac, util  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402
from tests import common, random_programs  # noqa: E402
from tests.emu import build_emu  # noqa: E402


def main():
  lo, hi = int(sys.argv[1]), int(sys.argv[2])
  hard = '--hard' in sys.argv
  start = time.time()
  ok = 0
  for seed in range(lo, hi):
    text, extent, kwargs = random_programs.program(seed, hard=hard)
    try:
      st = sodac.compile_source(text)
      inputs = random_programs.inputs_for(st, extent, seed)
      a = golden.run(st, inputs)
      want = emit_cpp.Oracle(st).run(inputs)
    except Exception as e:  # pylint: disable=broad-except
      print(seed, 'FRONT END / ORACLE', repr(e)[:300])
      print(text)
      continue
    index = common.box_index(st.valid_box('out', extent))
    if not np.array_equal(a['out'][index].view(np.uint8),
                          want['out'][index].view(np.uint8)):
      print(seed, 'ORACLES DIFFER')
      print(text)
      continue
    if not np.isfinite(a['out'][index].astype(np.float64)).all():
      print(seed, 'non-finite values')
    try:
      prog = launcher.CudaProgram(build_emu.build_emu_library(st, **kwargs))
    except util.SemanticError as e:
      print(seed, 'PLAN', e, kwargs)
      continue
    except Exception as e:  # pylint: disable=broad-except
      print(seed, 'BUILD FAILED', repr(e)[-600:])
      print(text)
      continue
    dtype = golden.np_dtype(st.output_stmts[0].haoda_type)
    outputs = {'out': np.full(extent[::-1], 77, dtype=dtype)}
    try:
      prog.run_host(inputs, outputs)
      common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)
      ok += 1
    except Exception as e:  # pylint: disable=broad-except
      print(seed, 'MISMATCH', repr(e)[:300], kwargs)
      print(text)
  print('ok %d of %d, %.0f s' % (ok, hi - lo, time.time() - start))


if __name__ == '__main__':
  main()
