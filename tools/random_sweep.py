#!/usr/bin/env python3
"""Differential sweep beyond the seeds the test-suite pins: random programs
(tests/random_programs.py) on cuda:0 against the g++ oracle.

  python tools/random_sweep.py build 40 120   # compile (no GPU needed)
  python tools/random_sweep.py run 40 120     # one JSON line per seed
  ... --hard: the generator's hard mode (wider windows, deeper DAGs, mixed
  stage types, let bindings, double / int64)
"""
import concurrent.futures
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from soda_b200 import sodac, util  # noqa: E402
from soda_b200.codegen.cuda import build as cuda_build  # noqa: E402
from soda_b200.codegen.cuda import launcher  # noqa: E402
from tests import common, random_programs  # noqa: E402


HARD = '--hard' in sys.argv


def build_one(seed):
  text, _, kwargs = random_programs.program(seed, hard=HARD)
  try:
    return cuda_build.build_library(sodac.compile_source(text),
                                    kwargs.get('time_block'),
                                    kwargs.get('options'))
  except util.SemanticError as e:
    return 'plan: %s' % e


def run_one(seed):
  from oracle import emit_cpp, golden
  text, extent, kwargs = random_programs.program(seed, hard=HARD)
  st = sodac.compile_source(text)
  inputs = random_programs.inputs_for(st, extent, seed)
  lib = build_one(seed)
  result = {'seed': seed, 'dim': st.dim, 'type': str(st.input_types[0]),
            'iterate': st.iterate, 'kwargs': kwargs}
  if lib.startswith('plan:'):
    result['status'] = lib
    return result
  try:
    prog = launcher.CudaProgram(lib)
    dtype = golden.np_dtype(st.output_stmts[0].haoda_type)
    outputs = {'out': np.full(extent[::-1], 77, dtype=dtype)}
    prog.run_host(inputs, outputs)
    common.assert_matches_oracle(st, extent, outputs,
                                 emit_cpp.Oracle(st).run(inputs), sentinel=77)
    result['status'] = 'ok'
  except Exception as e:  # pylint: disable=broad-except
    result['status'] = 'FAILED: %r' % (e,)
    result['program'] = text
  return result


def main():
  mode, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
  if mode == 'build':
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
      for seed, lib in zip(range(lo, hi), pool.map(build_one, range(lo, hi))):
        print(seed, os.path.basename(lib), flush=True)
  else:
    bad = 0
    for seed in range(lo, hi):
      result = run_one(seed)
      bad += result['status'] != 'ok'
      print(json.dumps(result), flush=True)
    print(json.dumps({'seeds': [lo, hi], 'not_ok': bad}), flush=True)


if __name__ == '__main__':
  main()
