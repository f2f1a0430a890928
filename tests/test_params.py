"""`param` statements through the whole stack.

The reference declares params as constant arrays next to the tensors
(src/soda/grammar.py ParamStmt), passes them to soda::app::<app> after the
tensors with the same (ptr, extent, stride, min) quadruple
(src/soda/codegen/frt/host.py:73-79), fills them with ``p[x][y] = x + y`` in
its test main (:530-543) and evaluates a reference ``p(i, j)`` as the constant
``p[i][j]`` (:580-586).  None of its tests/src programs uses one, so the two
programs here live in tests/src_extra."""
import os

import numpy as np
import pytest

from oracle import emit_cpp, golden
from soda_b200 import sodac, util
from soda_b200.codegen.cuda import emit, launcher, plan
from tests import common
from tests.emu import build_emu

EXTRA = os.path.join(common.ROOT, 'tests', 'src_extra')
CASES = [('conv2d_param', (300, 41), {}),
         ('conv2d_param', (200, 30), {'options': {'no_pack': True}}),
         ('scale3d_param', (150, 21, 9), {})]


def stencil(name, **overrides):
  with open(os.path.join(EXTRA, name + '.soda')) as fp:
    return sodac.compile_source(fp.read(), **overrides)


def test_front_end_round_trip_and_plan():
  st = stencil('conv2d_param')
  assert st.param_names == ('weight', 'bias')
  assert 'param float: weight[3][3]' in str(st)
  again = sodac.compile_source(str(st))
  assert str(again) == str(st)
  p = plan.make_tuned_pass_plan(st, 1)
  # params are constants, not loads: the DAG only has the tensors
  assert [n.name for n in p.nodes] == ['img', 'acc', 'out']
  text = emit.emit_program(st)
  assert 'SODA_CONSTANT float param_weight[9];' in text
  assert 'soda_gen::param_weight[5]' in text and 'soda_gen::param_bias[0]' in text
  # soda::app::<app> takes the params after the tensors
  assert text.index('var_out_ptr') < text.index('const float* var_weight_ptr')


def test_out_of_range_param_reference_is_rejected():
  text = open(os.path.join(EXTRA, 'conv2d_param.soda')).read().replace(
      'weight(2, 2)', 'weight(3, 0)')
  with pytest.raises(util.SemanticError):
    emit.emit_program(sodac.compile_source(text))


@pytest.mark.parametrize('name,extent,kwargs', CASES)
def test_oracles_agree_on_param_programs(name, extent, kwargs):
  st = stencil(name)
  inputs = common.make_inputs(st, extent, seed=2)
  for pattern in ('reference', 'random'):
    params = common.make_params(st, seed=3, pattern=pattern)
    a = golden.run(st, inputs, params=params)
    b = emit_cpp.Oracle(st).run(inputs, params=params)
    for out in st.output_names:
      index = common.box_index(st.valid_box(out, extent))
      assert np.array_equal(a[out][index].view(np.uint32),
                            b[out][index].view(np.uint32))


def _run(prog, st, extent, seed=5):
  inputs = common.make_inputs(st, extent, seed=seed)
  params = common.make_params(st, seed=seed)
  outputs = {n: np.full(extent[::-1], 77, dtype=d)
             for n, d in zip(prog.output_names, prog.output_dtypes)}
  prog.run_host(inputs, outputs, params=params)
  want = common.oracle_outputs(st, inputs, params=params)
  common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)
  # new values replace the old ones
  params2 = {k: v + 1 for k, v in params.items()}
  prog.run_host(inputs, outputs, params=params2)
  want2 = common.oracle_outputs(st, inputs, params=params2)
  common.assert_matches_oracle(st, extent, outputs, want2, sentinel=77)


@pytest.mark.parametrize('name,extent,kwargs', CASES)
def test_param_programs_under_emulation(name, extent, kwargs):
  st = stencil(name)
  small = tuple(max(8, e // 4) for e in extent)
  prog = launcher.CudaProgram(build_emu.build_emu_library(st, **kwargs))
  assert prog.param_names == list(st.param_names)
  _run(prog, st, small)


@pytest.mark.gpu
@pytest.mark.parametrize('name,extent,kwargs', CASES)
def test_param_programs_on_gpu(name, extent, kwargs):
  from soda_b200.codegen import cuda as cuda_backend
  st = stencil(name)
  prog = cuda_backend.compile_stencil(st, **kwargs)
  before = prog.launch_count()
  _run(prog, st, extent)
  assert prog.launch_count() > before
  with pytest.raises(ValueError):
    prog.run_host(common.make_inputs(st, extent))  # params missing


# ---- integer widths C++ does not have (second half) ----

def test_narrow_widths_are_lowered_to_containers_and_wraps():
  from soda_b200.optimization import widths
  st = stencil('narrow2d')
  assert widths.has_narrow_types(st)
  low = widths.lower(st)
  assert not widths.has_narrow_types(low)
  assert [str(t) for t in low.input_types + low.output_types] == ['uint8',
                                                                  'uint8']
  text = str(low)
  assert '& 63' in text and '^ 16) - 16' in text and '& 1023' in text
  # programs without such types pass through untouched
  plain = common.stencil('blur')
  assert widths.lower(plain) is plain
  # custom floats stay unsupported (the reference has no C type for them)
  with pytest.raises(util.SemanticError):
    emit.emit_program(sodac.compile_source(
        open(os.path.join(EXTRA, 'narrow2d.soda')).read().replace(
            'uint10', 'float18_3')))


def _narrow_inputs(st, extent, seed=9):
  rng = np.random.default_rng(seed)
  # container bytes with garbage above bit 5: an ap_uint<6> array cannot hold it
  return {'a': rng.integers(0, 256, extent[::-1]).astype(np.uint8)}


def test_narrow_widths_oracles_and_lowering_agree():
  from soda_b200.optimization import widths
  st = stencil('narrow2d')
  extent = (70, 19)
  inputs = _narrow_inputs(st, extent)
  a = golden.run(st, inputs)
  b = emit_cpp.Oracle(st).run(inputs)
  c = golden.run(widths.lower(st), inputs)  # the rewritten program
  index = common.box_index(st.valid_box('b', extent))
  assert np.array_equal(a['b'][index], b['b'][index])
  assert np.array_equal(a['b'][index], c['b'][index])
  assert a['b'][index].max() <= 63 and len(np.unique(a['b'][index])) > 20


def test_narrow_widths_under_emulation():
  st = stencil('narrow2d')
  extent = (90, 21)
  prog = launcher.CudaProgram(build_emu.build_emu_library(st))
  assert [str(d) for d in prog.input_dtypes + prog.output_dtypes] == ['uint8',
                                                                      'uint8']
  inputs = _narrow_inputs(st, extent)
  outputs = {'b': np.full(extent[::-1], 77, dtype=np.uint8)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs,
                               emit_cpp.Oracle(st).run(inputs), sentinel=77)


@pytest.mark.gpu
def test_narrow_widths_on_gpu():
  from soda_b200.codegen import cuda as cuda_backend
  st = stencil('narrow2d')
  extent = (1000, 211)
  prog = cuda_backend.compile_stencil(st, time_block=2)
  inputs = _narrow_inputs(st, extent)
  outputs = {'b': np.full(extent[::-1], 77, dtype=np.uint8)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs,
                               emit_cpp.Oracle(st).run(inputs), sentinel=77)


# ---- 33..63-bit integers: int64 containers ------------------------------------

def _wide_inputs(extent, seed=10):
  rng = np.random.default_rng(seed)
  # garbage above bit 39: an ap_uint<40> array cannot hold it
  return {'a': rng.integers(0, 2**62, extent[::-1]).astype(np.int64)}


def test_wide_widths_oracles_and_lowering_agree():
  from soda_b200.optimization import widths
  st = stencil('wide2d')
  low = widths.lower(st)
  assert not widths.has_narrow_types(low)
  assert [str(t) for t in low.input_types + low.output_types +
          tuple(low.local_types)] == ['int64'] * 4
  extent = (70, 19)
  inputs = _wide_inputs(extent)
  a = golden.run(st, inputs)
  b = emit_cpp.Oracle(st).run(inputs)
  c = golden.run(low, inputs)
  index = common.box_index(st.valid_box('b', extent))
  assert np.array_equal(a['b'][index], b['b'][index])
  assert np.array_equal(a['b'][index], c['b'][index])
  assert 2**39 < a['b'][index].max() < 2**40 and a['b'][index].min() >= 0
  assert len(np.unique(a['b'][index])) > 500


def test_wide_widths_under_emulation():
  st = stencil('wide2d')
  extent = (90, 21)
  prog = launcher.CudaProgram(build_emu.build_emu_library(st))
  assert [str(d) for d in prog.input_dtypes + prog.output_dtypes] == ['int64',
                                                                      'int64']
  inputs = _wide_inputs(extent)
  outputs = {'b': np.full(extent[::-1], 77, dtype=np.int64)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs,
                               emit_cpp.Oracle(st).run(inputs), sentinel=77)


@pytest.mark.gpu
def test_wide_widths_on_gpu():
  from soda_b200.codegen import cuda as cuda_backend
  st = stencil('wide2d')
  extent = (1000, 211)
  prog = cuda_backend.compile_stencil(st)
  inputs = _wide_inputs(extent)
  outputs = {'b': np.full(extent[::-1], 77, dtype=np.int64)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs,
                               emit_cpp.Oracle(st).run(inputs), sentinel=77)


def test_narrow_widths_with_let_bindings():
  """Found by the random-program sweep: the let lines of a multi-line
  statement must not leak into the header of the lowered program."""
  from soda_b200.optimization import widths
  st = sodac.compile_source('''kernel: narrowlet
burst width: 64
unroll factor: 2
iterate: 2
input uint6: a(32, *)
local uint6:
  int32 t = (a(0, 0) + a(1, -1)) / 2
  u = t - a(0, 1)
  m(0, 0) = (t & 255) - (u % 7) + a(-1, 0)
output uint6: b(0, 0) = (m(0, 0) + m(1, 1) + a(0, 0)) * 3 % 17
''')
  low = widths.lower(st)
  assert not widths.has_narrow_types(low)
  assert [str(l.name) for l in low.local_stmts[0].let] == ['t', 'u']
  extent = (70, 19)
  inputs = _narrow_inputs(st, extent)
  a = golden.run(st, inputs)
  b = emit_cpp.Oracle(st).run(inputs)
  c = golden.run(low, inputs)
  index = common.box_index(st.valid_box('b', extent))
  assert np.array_equal(a['b'][index], b['b'][index])
  assert np.array_equal(a['b'][index], c['b'][index])
  prog = launcher.CudaProgram(build_emu.build_emu_library(st))
  outputs = {'b': np.full(extent[::-1], 77, dtype=np.uint8)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs, b, sentinel=77)
