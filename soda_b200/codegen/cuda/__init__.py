"""``soda.codegen.cuda``: the B200 backend of ``sodac``.

Follows the reference's backend plugin convention - a module exporting
``add_arguments(argument_group)`` and ``print_code(stencil, args)`` that
``sodac`` registers by hand (reference: src/soda/sodac.py:99-102,198-200;
e.g. the FRT host backend, src/soda/codegen/frt/core.py:10-27).  ``-`` as a
file name means stdout, as there.

Outputs (each optional):
  --cuda-kernel FILE   the generated translation unit (functors + plan tables)
  --cuda-lib FILE.so   the compiled program library (C ABI: include/soda_cuda.h)
  --cuda-host FILE.py  a small Python host script that loads the library

``compile_stencil`` / ``load`` are the programmatic equivalents.
"""
import argparse
import shutil
import sys
from typing import Dict, Optional

from soda_b200.codegen.cuda import build, emit, launcher, plan

__all__ = [
    'add_arguments', 'print_code', 'compile_stencil', 'load', 'options_from_args'
]


def add_arguments(parser) -> None:
  parser.add_argument('--cuda-kernel', type=str, dest='cuda_kernel',
                      metavar='file',
                      help='generated CUDA translation unit (functors and plan '
                      'tables instantiating the hand-written templates)')
  parser.add_argument('--cuda-lib', type=str, dest='cuda_lib', metavar='file',
                      help='compiled program library for sm_100a')
  parser.add_argument('--cuda-host', type=str, dest='cuda_host',
                      metavar='file', help='Python host script for the library')
  parser.add_argument('--cuda-time-block', type=int, dest='cuda_time_block',
                      metavar='N',
                      help='iterations fused per HBM round trip (default: up '
                      'to 4 in 2-D, 2 in 3-D)')
  parser.add_argument('--cuda-cells', type=int, dest='cuda_cells', metavar='N',
                      help='cells per lane in dimension 0 (default: 16 bytes '
                      'worth, 32 for light 2-D fp32 programs)')
  parser.add_argument('--cuda-rows', type=int, dest='cuda_rows', metavar='N',
                      help='3-D tile height (default: 4-6 x the dimension-1 '
                      'halo, at least 8)')
  parser.add_argument('--cuda-patch-rows', type=int, dest='cuda_patch_rows',
                      metavar='N',
                      help='3-D tile rows owned by one thread (default: 4, 2 '
                      'or 1 by register budget)')
  parser.add_argument('--cuda-warps', type=int, dest='cuda_warps', metavar='N',
                      help='2-D strips (warps) per CTA (default 2 or 4)')
  parser.add_argument('--cuda-min-blocks', type=int, dest='cuda_min_blocks',
                      metavar='N', help='CTAs per SM to ask the compiler for')
  parser.add_argument('--cuda-no-pipeline', action='store_true',
                      dest='cuda_no_pipeline',
                      help='2-D: evaluate producers before consumers within a '
                      'step instead of the pipelined DAG schedule')
  parser.add_argument('--cuda-chunk', type=int, dest='cuda_chunk', metavar='N',
                      help='2-D rows per TMA box')
  parser.add_argument('--cuda-stages', type=int, dest='cuda_stages',
                      metavar='N', help='2-D TMA ring slots per warp')
  parser.add_argument('--cuda-no-pack', action='store_true',
                      dest='cuda_no_pack',
                      help='do not use packed fp32 pairs (FADD2/FMUL2) for '
                      'float add/mul programs')
  parser.add_argument('--cuda-gpus', type=int, dest='cuda_gpus', metavar='N',
                      help='the program-named entry point splits the grid '
                      'along the outermost dimension over N devices of the '
                      'calling process (soda_cuda_multi_run_host); a caller '
                      'can still override it per call')
  parser.add_argument('--cuda-fast-fp', action='store_true',
                      dest='cuda_fast_fp',
                      help='allow FMA contraction (default: off, results are '
                      'bit-identical to g++ without -ffp-contract)')


  parser.add_argument('--cuda-pow2-fma', action='store_true',
                      dest='cuda_pow2_fma',
                      help='fuse `c * x + y` when the float literal c is a '
                      'power of two: the product is exact, so the result is '
                      'the un-fused one unless c * x is subnormal')


def options_from_args(args: Optional[argparse.Namespace]) -> Dict:
  get = lambda name: getattr(args, name, None) if args is not None else None
  options = {
      'cells': get('cuda_cells'),
      'rows': get('cuda_rows'),
      'warps': get('cuda_warps'),
      'chunk': get('cuda_chunk'),
      'stages': get('cuda_stages'),
      'cy': get('cuda_patch_rows'),
      'min_blocks': get('cuda_min_blocks'),
      'gpus': get('cuda_gpus'),
      'no_pipeline': bool(get('cuda_no_pipeline')),
      'fast_fp': bool(get('cuda_fast_fp')),
      'no_pack': bool(get('cuda_no_pack')),
      'pow2_fma': bool(get('cuda_pow2_fma')),
  }
  return {k: v for k, v in options.items() if v}


HOST_TEMPLATE = '''#!/usr/bin/env python3
"""Host for the SODA program `{app}` compiled for NVIDIA B200 by sodac."""
import sys

import numpy as np

from soda_b200.codegen.cuda import launcher

program = launcher.CudaProgram({lib!r})


def run(*arrays):
  """arrays: one NumPy array per input ({inputs}); returns the outputs."""
  return program.run_host(dict(zip(program.input_names, arrays)))


if __name__ == '__main__':
  extent = [int(x) for x in sys.argv[1:]] or {default_extent!r}
  rng = np.random.default_rng(0)
  arrays = []
  for dtype in program.input_dtypes:
    if dtype.kind == 'f':
      arrays.append(rng.random(extent[::-1]).astype(dtype))
    else:
      arrays.append(np.indices(extent[::-1]).sum(axis=0).astype(dtype))
  outputs = run(*arrays)
  for name, array in outputs.items():
    print(name, array.shape, array.dtype, 'checksum', float(array.sum()))
'''


def _write(path: str, text: str) -> None:
  if path == '-':
    sys.stdout.write(text)
  else:
    with open(path, 'w') as fp:
      fp.write(text)


def print_code(stencil, args: argparse.Namespace) -> None:
  kernel = getattr(args, 'cuda_kernel', None)
  lib = getattr(args, 'cuda_lib', None)
  host = getattr(args, 'cuda_host', None)
  if kernel is None and lib is None and host is None:
    return
  options = options_from_args(args)
  time_block = getattr(args, 'cuda_time_block', None)
  if kernel is not None:
    _write(kernel, emit.emit_program(stencil, time_block, options))
  lib_path = None
  if lib is not None:
    lib_path = build.build_library(stencil, time_block, options, output=lib)
  if host is not None:
    if lib_path is None:
      lib_path = build.build_library(stencil, time_block, options)
    default_extent = [
        s if s else 64 for s in stencil.tile_size
    ]
    _write(
        host,
        HOST_TEMPLATE.format(app=stencil.app_name,
                             lib=lib_path,
                             inputs=', '.join(stencil.input_names),
                             default_extent=default_extent))


def compile_stencil(stencil,
                    time_block: Optional[int] = None,
                    options: Optional[Dict] = None) -> 'launcher.CudaProgram':
  """Stencil -> loaded program (built into soda_b200/_build, cached)."""
  return launcher.CudaProgram(build.build_library(stencil, time_block, options))


def load(lib_path: str) -> 'launcher.CudaProgram':
  return launcher.CudaProgram(lib_path)
