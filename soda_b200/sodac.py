"""``sodac``: the SODA compiler driver, with the CUDA (B200) target.

Same command line as the reference for everything that reaches the stencil IR
(reference: src/soda/sodac.py:26-116,153-194): the DSL directives
``burst width``, ``unroll factor``, ``tile size``, ``iterate``, ``border``,
``cluster`` and the ``dram`` placement can be overridden from the CLI, the
optimisation group adds ``--inline`` and ``--computation-reuse``, and each
backend registers its own argument group through the plugin convention
``add_arguments(group)`` / ``print_code(stencil, args)`` (reference:
src/soda/sodac.py:99-102,198-200).  The FPGA backends of the reference are out
of scope here; the CUDA backend takes their place.
"""
import argparse
import logging
import sys
from typing import List, Optional

from soda_b200 import core, grammar, util
from soda_b200.codegen import cuda as cuda_backend
from soda_b200.optimization import args as opt_args

logger = logging.getLogger('sodac')


def build_parser() -> argparse.ArgumentParser:
  parser = argparse.ArgumentParser(
      prog='sodac',
      description='Stencil with Optimized Dataflow Architecture (SODA) '
      'compiler, B200 CUDA target')
  parser.add_argument('--verbose', '-v', action='count', dest='verbose',
                      help='increase verbosity')
  parser.add_argument('--quiet', '-q', action='count', dest='quiet',
                      help='decrease verbosity')
  parser.add_argument('--recursion-limit', type=int, dest='recursion_limit',
                      help='override Python recursion limit')
  parser.add_argument('--burst-width', type=int, dest='burst_width',
                      help='override burst width')
  parser.add_argument('--unroll-factor', type=int, metavar='UNROLL_FACTOR',
                      dest='unroll_factor', help='override unroll factor')
  parser.add_argument('--replication-factor', type=int,
                      metavar='REPLICATION_FACTOR', dest='replication_factor',
                      help='override replication factor')
  parser.add_argument('--tile-size', type=int, nargs='+', metavar='TILE_SIZE',
                      dest='tile_size',
                      help='override tile size; 0 means no overriding on that '
                      'dimension')
  parser.add_argument('--dram-in', type=str, dest='dram_in',
                      help='override DRAM configuration for input')
  parser.add_argument('--dram-out', type=str, dest='dram_out',
                      help='override DRAM configuration for output')
  parser.add_argument('--iterate', type=int, metavar='#ITERATION',
                      dest='iterate',
                      help='override iterate directive; repeat execution '
                      'multiple times iteratively')
  parser.add_argument('--border', type=str, metavar='(ignore|preserve)',
                      dest='border', help='override border handling strategy')
  parser.add_argument('--cluster', type=str,
                      metavar='(none|fine|coarse|full)', dest='cluster',
                      help='module clustering level (accepted for '
                      'compatibility; the CUDA backend always fuses '
                      'everything)')
  parser.add_argument('--math-precision', type=str, metavar='(double|float)',
                      dest='math_precision',
                      help='how sqrt / fabs / floor / ceil of a float are '
                      'evaluated: through double (default: what g++ does with '
                      'the generated code and plain <cmath>) or in float (the '
                      'std:: float overloads in scope, as Xilinx headers may '
                      'arrange); applies to every backend and to the oracle')
  parser.add_argument(type=str, dest='soda_src', metavar='file',
                      help='soda source code, - for stdin')
  cuda_backend.add_arguments(parser.add_argument_group('CUDA (B200) backend'))
  opt_args.add_arguments(parser.add_argument_group('SODA optimizations'))
  return parser


def stencil_from_program(program: grammar.SodaProgram,
                         args: Optional[argparse.Namespace] = None
                        ) -> core.Stencil:
  """Applies CLI overrides and builds the Stencil
  (reference: src/soda/sodac.py:153-194)."""
  get = (lambda name: getattr(args, name, None)) if args is not None else (
      lambda name: None)

  tile_size = []
  override = get('tile_size')
  for dim in range(program.dim - 1):
    if override is not None and dim < len(override) and override[dim] > 0:
      tile_size.append(override[dim])
    else:
      tile_size.append(program.tile_size[dim])
  tile_size.append(0)

  if get('replication_factor') is None:
    unroll_factor = get('unroll_factor')
    if unroll_factor is None:
      unroll_factor = program.unroll_factor
    replication_factor = 1
  else:
    unroll_factor = replication_factor = get('replication_factor')

  def pick(name):
    value = get(name)
    return value if value is not None else getattr(program, name)

  return core.Stencil(
      burst_width=pick('burst_width'),
      border=pick('border'),
      iterate=pick('iterate'),
      cluster=pick('cluster'),
      dram_in=get('dram_in'),
      dram_out=get('dram_out'),
      app_name=program.app_name,
      input_stmts=program.input_stmts,
      param_stmts=program.param_stmts,
      local_stmts=program.local_stmts,
      output_stmts=program.output_stmts,
      dim=program.dim,
      tile_size=tile_size,
      unroll_factor=unroll_factor,
      replication_factor=replication_factor,
      math_precision=get('math_precision'),
      optimizations=opt_args.get_kwargs(args) if args is not None and hasattr(
          args, 'computation_reuse') else {},
  )


def compile_source(text: str, **overrides) -> core.Stencil:
  """Programmatic front door: SODA text -> Stencil, with optional overrides
  named like the CLI destinations (iterate=, tile_size=, inline=, ...)."""
  ns = argparse.Namespace(inline='no', computation_reuse='no')
  for key, value in overrides.items():
    setattr(ns, key, value)
  return stencil_from_program(grammar.parse(text), ns)


def main(argv: Optional[List[str]] = None) -> int:
  parser = build_parser()
  args = parser.parse_args(sys.argv[1:] if argv is None else argv)
  level = logging.WARNING + 10 * ((args.quiet or 0) - (args.verbose or 0))
  logging.basicConfig(
      level=min(max(level, logging.DEBUG), logging.CRITICAL),
      format='%(levelname)s:%(name)s:%(lineno)d: %(message)s')
  if args.recursion_limit is not None and \
      args.recursion_limit > sys.getrecursionlimit():
    sys.setrecursionlimit(args.recursion_limit)

  try:
    if args.soda_src == '-':
      text = sys.stdin.read()
    else:
      with open(args.soda_src) as soda_file:
        text = soda_file.read()
    program = grammar.parse(text)
    stencil = stencil_from_program(program, args)
    cuda_backend.print_code(stencil, args)
  except util.SemanticError as e:  # includes syntax errors
    logger.error(e)
    return 1
  except util.SemanticWarn as w:
    logger.warning(w)
  return 0


if __name__ == '__main__':
  sys.exit(main())
