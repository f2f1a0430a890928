"""--inline, restating src/tests/optimization/test_inline.py:24-113."""
from soda_b200 import core, grammar
from soda_b200.optimization import inline

HEADER = '''
kernel: blur
burst width: 512
unroll factor: 16
input float: t0(233, *)
%s
iterate: 1
border: preserve
cluster: none
'''


def make(body):
  program = grammar.parse(HEADER % body)
  stencil = core.Stencil(**{**program.__dict__, 'replication_factor': 1})
  inline.inline(stencil)
  return stencil


def test_simple_inlining():
  stencil = make('local float: t1(-1, -2) = t0(0, 1)\n'
                 'output float: t2(4, 2) = t1(2, 3)')
  assert len(stencil.local_stmts) == 0
  assert str(stencil.output_stmts[0]) == \
      'output dram 0 float: t2(4, 2) = t0(3, 6)'


def test_let_in_local():
  stencil = make('local float: float l = t0(0, 1) t1(-1, -2) = l\n'
                 'output float: t2(4, 2) = t1(2, 3)')
  assert len(stencil.local_stmts) == 0
  assert str(stencil.output_stmts[0]) == (
      'output dram 0 float:\n  float l = t0(3, 6)\n  t2(4, 2) = l')


def test_let_in_output():
  stencil = make('local float: t1(-1, -2) = t0(0, 1)\n'
                 'output float: float l = t1(2, 3) t2(4, 2) = l')
  assert len(stencil.local_stmts) == 0
  assert str(stencil.output_stmts[0]) == (
      'output dram 0 float:\n  float l = t0(3, 6)\n  t2(4, 2) = l')


def test_access_in_different_stmts():
  stencil = make('local float: t1(-1, -2) = t0(0, 1)\n'
                 'local float: t2(0, 0) = t1(0, 0)\n'
                 'output float: t3(4, 2) = t2(0, 0) + t1(0, 0) + t2(0, 1)')
  assert len(stencil.local_stmts) == 2
  assert str(stencil.output_stmts[0]) == (
      'output dram 0 float: t3(4, 2) = t2(0, 0) + t1(0, 0) + t2(0, 1)')


def test_inline_flag_keeps_results():
  """--inline must not change what the program computes."""
  import numpy as np
  from oracle import golden
  from soda_b200 import sodac
  from tests import common
  plain = common.stencil('blur')
  inlined = sodac.compile_source(common.source('blur'), inline='yes')
  assert len(inlined.local_stmts) == 1  # blur_x is loaded three times
  src = common.source('sobel2d').replace(
      'mag(0, 0) = 65535 - (mag_x(0, 0) * mag_x(0, 0) + mag_y(0, 0) * '
      'mag_y(0, 0))', 'mag(0, 0) = 65535 - (mag_x(0, 0) + mag_y(0, 0))')
  a = sodac.compile_source(src)
  b = sodac.compile_source(src, inline='yes')
  assert len(a.local_stmts) == 2 and len(b.local_stmts) == 0
  extent = (40, 9)
  inputs = common.make_inputs(a, extent, seed=3)
  ra, rb = golden.run(a, inputs), golden.run(b, inputs)
  inside = common.box_index(a.valid_box('mag', extent))
  assert np.array_equal(ra['mag'][inside], rb['mag'][inside])
