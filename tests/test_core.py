"""Stencil IR tests, restating src/tests/test_core.py of the reference (error
messages and the iterate chain) and pinning the window analyses against the
values the reference's own functions produce (tests/golden/windows.json, made
by tests/golden/make_reference_fixtures.py from /root/reference)."""
import copy
import json
import os

import pytest

from soda_b200 import core, grammar, ir, sodac, util
from tests import common

BLUR = r'''
kernel: blur
burst width: 512
unroll factor: 16
input uint16: input(2000, *)
local uint16: tmp(0,0)=(input(-1,0)+input(0,0)+input(1,0))/3
output uint16: output(0,0)=(tmp(0,-1)+tmp(0,0)+tmp(0,1))/3
iterate: 2
border: preserve
cluster: none
'''


def blur_args():
  program = grammar.parse(BLUR)
  return {**program.__dict__, 'replication_factor': 1}


def test_number_of_inputs_differs_from_outputs():
  args = blur_args()
  extra = grammar.InputStmt(haoda_type=ir.Type('uint16'), name='bar',
                            tile_size=[233], dram=())
  args['input_stmts'] = args['input_stmts'] + [extra]
  with pytest.raises(util.SemanticError) as err:
    core.Stencil(**args)
  assert str(err.value) == (
      'number of input tensors must be the same as output if iterate > 1 '
      'times, currently there are 2 input(s) but 1 output(s)')


def test_input_type_differs_from_output():
  args = blur_args()
  stmt = copy.copy(args['input_stmts'][0])
  stmt.haoda_type = ir.Type('half')
  args['input_stmts'] = [stmt]
  with pytest.raises(util.SemanticError) as err:
    core.Stencil(**args)
  assert str(err.value) == (
      'input must have the same type(s) as output if iterate > 1 times, '
      'current input has type [half] but output has type [uint16]')


def test_iterate_must_be_positive():
  args = blur_args()
  args['iterate'] = 0
  with pytest.raises(util.SemanticError) as err:
    core.Stencil(**args)
  assert str(err.value) == 'cannot iterate 0 times'


def test_high_level_dag_construction():
  stencil = core.Stencil(**blur_args())
  names = ('input', 'tmp', 'input_iter1', 'tmp_iter1', 'output')
  assert tuple(stencil.tensors) == names
  assert tuple(t.name for t in stencil.chronological_tensors) == names
  assert stencil.tensors['input_iter1'].parents.keys() == {'tmp'}
  assert stencil.tensors['tmp_iter1'].parents.keys() == {'input_iter1'}


def test_doc_example_distance():
  """docs/data-layout.md:18-25: centred 3x3 window on a 100-wide tile has
  stencil distance 202."""
  stencil = sodac.compile_source(
      'kernel: k\nburst width: 64\nunroll factor: 1\niterate: 1\n'
      'input float: a(100, *)\n'
      'output float: b(0, 0) = a(-1, -1) + a(0, -1) + a(1, -1) + a(-1, 0) + '
      'a(0, 0) + a(1, 0) + a(-1, 1) + a(0, 1) + a(1, 1)')
  assert stencil.stencil_distance == 202
  assert core.get_stencil_dim(stencil.stencil_window) == [3, 3]


# SURVEY.md appendix B (computed there with the reference's own functions)
APPENDIX_B = {
    'blur': (9, (0, 0), (2, 2), 4002),
    'contrast': (197, (0, 0), (16, 16), 7688),
    'denoise2d': (13, (-2, -2), (2, 2), 130),
    'denoise3d': (25, (-2, -2, -2), (2, 2, 2), 4162),
    'erosion': (361, (-9, -9), (9, 9), 8658),
    'heat3d': (25, (-2, -2, -2), (2, 2, 2), 4162),
    'jacobi2d': (13, (-2, -2), (2, 2), 130),
    'jacobi3d': (25, (-2, -2, -2), (2, 2, 2), 4162),
    'seidel2d': (25, (-2, -2), (2, 2), 132),
    'sobel2d': (8, (-1, -1), (1, 1), 66),
    'xcorr': (361, (-9, -9), (9, 9), 8658),
}


@pytest.mark.parametrize('name', common.PROGRAMS)
def test_window_analyses(name):
  stencil = common.stencil(name)
  points, lo, hi, distance = APPENDIX_B[name]
  assert len(stencil.stencil_window) == points
  assert stencil.window_bounds[stencil.output_names[0]] == (lo, hi)
  assert stencil.stencil_distance == distance


def test_window_bounds_match_enumerated_window():
  """The closed-form bounding boxes equal the bounding box of the enumerated
  dependency cone for every tensor of every program."""
  for name in common.PROGRAMS:
    stencil = common.stencil(name)
    inputs = [stencil.tensors[n] for n in stencil.input_names]
    for tensor in stencil.chronological_tensors:
      if tensor.is_input():
        continue
      window = core.get_overall_stencil_window(inputs, tensor)
      lo = tuple(min(p[d] for p in window) for d in range(stencil.dim))
      hi = tuple(max(p[d] for p in window) for d in range(stencil.dim))
      assert stencil.window_bounds[tensor.name] == (lo, hi), tensor.name


def test_reference_fixture_windows():
  path = os.path.join(os.path.dirname(__file__), 'golden', 'windows.json')
  if not os.path.exists(path):
    pytest.skip('fixture not generated')
  with open(path) as fp:
    fixtures = json.load(fp)
  for case in fixtures['cases']:
    stencil = common.stencil(case['program'], **case['overrides'])
    inputs = [stencil.tensors[n] for n in stencil.input_names]
    for tname, want in case['tensors'].items():
      window = core.get_overall_stencil_window(inputs, stencil.tensors[tname])
      assert [list(p) for p in window] == want['window'], (case, tname)
      assert core.get_stencil_distance(window, stencil.tile_size) == \
          want['distance']
      assert list(core.get_stencil_window_offset(window)) == want['offset']
      assert core.get_stencil_dim(window) == want['dim']
    assert stencil.stencil_distance == case['stencil_distance']


def test_reuse_buffers_jacobi_tile_2000():
  """README.md:127-152 of the reference: 5-point Jacobi at tile 2000 touches
  stream offsets {0, 1999, 2000, 2001, 4000}, i.e. a 2-line reuse buffer."""
  stencil = sodac.compile_source(common.source('jacobi2d'), iterate=1,
                                 tile_size=[2000], unroll_factor=1)
  chains = core._get_reuse_chains(stencil.tile_size, stencil.tensors['t1'], 1)
  assert chains == [(0, 1999, 2000, 2001, 4000)]
  buffer = stencil.reuse_buffers['t1']
  assert buffer[0] == 4001
  assert sum(stencil.reuse_buffer_lengths['t1'].values()) == 4000


def test_rebalance_contrast():
  """src/soda/optimization/inline.py:175-262: the 197-term float sum is cut
  into 7 groups of at most 32 terms; 6 become cr_var locals."""
  stencil = common.stencil('contrast')
  assert [s.name for s in stencil.local_stmts] == [
      'cr_var_%d' % i for i in range(6)
  ]
  for stmt in stencil.local_stmts:
    assert len(ir.unwrap(stmt.expr).operand) == 32
  out = ir.unwrap(stencil.output_stmts[0].expr)
  assert len(out.operand) == 5 + 6
  assert [str(o) for o in out.operand[-6:]] == [
      'cr_var_%d(0, 0)' % i for i in range(6)
  ]


def test_cli_overrides():
  stencil = sodac.compile_source(common.source('blur'), iterate=2,
                                 tile_size=[1000], unroll_factor=4,
                                 burst_width=128, border='preserve')
  assert (stencil.iterate, stencil.tile_size, stencil.unroll_factor,
          stencil.burst_width, stencil.border) == (2, (1000, 0), 4, 128,
                                                   'preserve')
  assert tuple(stencil.tensors) == ('input', 'blur_x', 'input_iter1',
                                    'blur_x_iter1', 'blur_y')
