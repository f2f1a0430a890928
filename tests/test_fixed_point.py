"""Fixed-point types ``(u)intN_M`` (the reference prints ``ap_(u)fixed<N, M>``;
its tests only parse and print them, src/tests/test_grammar.py:40-54):
soda_b200/optimization/fixed_point.py rewrites them to scaled integers for the
CUDA backend and the g++ oracle, oracle/golden.py evaluates them natively.
Hand-computed values pin the rules (exact + - *, AP_TRN stores, AP_WRAP);
the evaluators must then agree bit for bit on a whole program."""
import os

import numpy as np
import pytest

from oracle import emit_cpp, golden
from soda_b200 import ir, sodac, util
from soda_b200.codegen.cuda import emit, launcher
from soda_b200.optimization import fixed_point, widths
from tests import common
from tests.emu import build_emu

EXTRA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'src_extra')


def stencil(name, **overrides):
  with open(os.path.join(EXTRA, name + '.soda')) as fp:
    return sodac.compile_source(fp.read(), **overrides)


def test_type_properties():
  t = ir.Type('uint18_3')
  assert t.is_fixed and not t.is_float and not t.is_signed
  assert (t.width_in_bits, t.frac_bits) == (18, 15)
  assert str(t.raw_type) == 'uint18' and str(t.container) == 'uint32'
  t = ir.Type('int40_20')
  assert t.is_signed and t.frac_bits == 20 and str(t.container) == 'int64'
  assert ir.Type('int32').frac_bits == 0 and not ir.Type('float18_3').is_fixed
  assert ir.Type.exact_fixed(11).frac_bits == 11


POINTWISE = '''kernel: fx
burst width: 64
unroll factor: 2
iterate: 1
input uint8_4: a(32, *)
input int8_4: c(32, *)
output int8_6: lo(0, 0) = c(0, 0)
output uint8_4: sq(0, 0) = a(0, 0) * a(0, 0)
output int8_4: neg(0, 0) = c(0, 0) - a(0, 0) * 2
output uint8_4: fl(0, 0) = uint8_4(float(a(0, 0)) * 0.3f)
output int16: whole(0, 0) = int16(c(0, 0) * 3)
output float: real(0, 0) = float(c(0, 0))
output uint8: cmp(0, 0) = uint8(c(0, 0) < a(0, 0)) + uint8(max(a(0, 0), c(0, 0)) == a(0, 0)) * 2
'''


def test_hand_computed_values():
  st = sodac.compile_source(POINTWISE)
  shape = (3, 32)
  a = np.zeros(shape, np.uint8)
  c = np.zeros(shape, np.int8)
  # a = 1.5 (raw 24), 15.9375 (raw 255), 0.0625 (raw 1)
  a[0, :3] = [24, 255, 1]
  # c = -1.0625 (raw -17), 7.9375 (raw 127), -8.0 (raw -128)
  c[0, :3] = [-17, 127, -128]
  want = {
      # int8_6 has 2 fractional bits; toward minus infinity: floor(-17 / 4) = -5 = -1.25
      'lo': [-5, 31, -32],
      # 1.5^2 = 2.25 = raw 36; 15.9375^2 = 254.00390625 -> raw 4064.06 -> 4064
      # & 255 = 224; 0.0625^2 -> raw floor(1 / 16) = 0
      'sq': [36, 224, 0],
      # -1.0625 - 3 = -4.0625 = raw -65; 7.9375 - 31.875 = raw -383 -> wrap to
      # 8 bits signed: -383 & 255 = 129 -> -127; -8 - 0.125 = raw -130 -> 126
      'neg': [-65, -127, 126],
      # floor(1.5f * 0.3f * 16) = floor(7.2) = 7; 15.9375 * 0.3 * 16 = 76.5 ->
      # 76; 0.0625 * 0.3 * 16 = 0.3 -> 0
      'fl': [7, 76, 0],
      # toward zero: -3.1875 -> -3; 23.8125 -> 23; -24 -> -24
      'whole': [-3, 23, -24],
      'real': [-1.0625, 7.9375, -8.0],
      # c < a: 1, 1, 1; max(a, c) == a: 1, 1, 1
      'cmp': [3, 3, 3],
  }
  inputs = {'a': a, 'c': c}
  for result in (golden.run(st, inputs), emit_cpp.Oracle(st).run(inputs),
                 golden.run(widths.lower(fixed_point.lower(st)), inputs)):
    for name, values in want.items():
      assert list(result[name][0, :3]) == values, name


def test_lowering_removes_fixed_types_and_binds_shifts_to_lets():
  st = stencil('fixed2d')
  assert fixed_point.has_fixed_types(st)
  low = fixed_point.lower(st)
  assert not fixed_point.has_fixed_types(low)
  assert [str(t) for t in low.input_types + low.output_types] == ['uint18',
                                                                  'uint18']
  text = fixed_point.lower_text(str(st))
  assert 'int64 fx0 = ' in text and '(fx0 - (fx0 & 1)) / 2' in text
  plain = common.stencil('blur')
  assert fixed_point.lower(plain) is plain


def _inputs(extent, seed=5):
  rng = np.random.default_rng(seed)
  # garbage above bit 17: an ap_ufixed<18, 3> array cannot hold it
  return {'a': rng.integers(0, 2**31, extent[::-1]).astype(np.uint32)}


def test_evaluators_agree():
  st = stencil('fixed2d')
  extent = (70, 19)
  inputs = _inputs(extent)
  native = golden.run(st, inputs)
  rewritten = golden.run(widths.lower(fixed_point.lower(st)), inputs)
  compiled = emit_cpp.Oracle(st).run(inputs)
  index = common.box_index(st.valid_box('b', extent))
  assert np.array_equal(native['b'][index], rewritten['b'][index])
  assert np.array_equal(native['b'][index], compiled['b'][index])
  assert native['b'][index].max() < 2**18
  assert len(np.unique(native['b'][index])) > 500


def test_under_emulation():
  st = stencil('fixed2d')
  extent = (90, 21)
  prog = launcher.CudaProgram(build_emu.build_emu_library(st))
  assert [str(d) for d in prog.input_dtypes + prog.output_dtypes] == ['uint32',
                                                                      'uint32']
  inputs = _inputs(extent)
  outputs = {'b': np.full(extent[::-1], 77, dtype=np.uint32)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs, golden.run(st, inputs),
                               sentinel=77)


@pytest.mark.gpu
def test_on_gpu():
  from soda_b200.codegen import cuda as cuda_backend
  st = stencil('fixed2d')
  extent = (1000, 211)
  prog = cuda_backend.compile_stencil(st)
  inputs = _inputs(extent)
  outputs = {'b': np.full(extent[::-1], 77, dtype=np.uint32)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs, golden.run(st, inputs),
                               sentinel=77)


@pytest.mark.parametrize('expr,message', [
    ('a(0, 0) / a(1, 0)', 'not supported on fixed-point'),
    ('a(0, 0) % 3', 'not supported on fixed-point'),
    ('a(0, 0) * 0.5f', 'mixes fixed-point and floating-point'),
    ('sqrt(a(0, 0))', 'cast to float first'),
    ('a(0, 0) & 3', 'not supported on fixed-point'),
])
def test_inexact_operations_are_rejected(expr, message):
  st = sodac.compile_source('''kernel: fxbad
burst width: 64
unroll factor: 2
iterate: 1
input uint8_4: a(32, *)
output uint8_4: b(0, 0) = %s
''' % expr)
  with pytest.raises(util.SemanticError, match=message):
    emit.emit_program(st)


def test_custom_floats_stay_rejected():
  st = sodac.compile_source('''kernel: fl
burst width: 64
unroll factor: 2
iterate: 1
input float18_3: a(32, *)
output float18_3: b(0, 0) = a(0, 0) + a(1, 0)
''')
  with pytest.raises(util.SemanticError):
    emit.emit_program(st)


def _random_fixed_program(seed):
  """A seeded random fixed-point program: two stages over one input, types
  with random widths / integer bits, sums, differences and products of taps,
  integer literals, casts to other fixed-point types, comparisons."""
  rng = np.random.default_rng(7000 + seed)

  def fixed_type(max_bits=18):
    bits = int(rng.integers(6, max_bits + 1))
    integer = int(rng.integers(1, bits))
    return '%sint%d_%d' % ('u' if rng.random() < 0.5 else '', bits, integer)

  def tap(name, store=(0, 0)):
    return '%s(%d, %d)' % (name, store[0] + int(rng.integers(-1, 2)),
                           store[1] + int(rng.integers(-1, 2)))

  def expr(name):
    terms = ['%s(0, 0)' % name]
    for _ in range(int(rng.integers(1, 4))):
      kind = rng.integers(5)
      t = tap(name)
      if kind == 0:
        terms.append(t)
      elif kind == 1:
        terms.append('%s * %d' % (t, int(rng.integers(2, 6))))
      elif kind == 2:
        terms.append('%s * %s' % (t, tap(name)))
      elif kind == 3:
        terms.append('%s(%s)' % (fixed_type(12), t))
      else:
        terms.append('max(%s, %s)' % (t, tap(name)))
    text = terms[0]
    for term in terms[1:]:
      text += (' + ' if rng.random() < 0.6 else ' - ') + term
    if rng.random() < 0.3:
      text = '(%s) * (%s > %s)' % (text, tap(name), tap(name))
    return text

  in_t, mid_t, out_t = fixed_type(16), fixed_type(), fixed_type()
  return '''kernel: fxrnd%d
burst width: 64
unroll factor: 2
iterate: %d
input %s: a(32, *)
local %s: m(0, 0) = %s
output %s: b(0, 0) = %s
''' % (seed, 1 if in_t != out_t else int(rng.integers(1, 3)), in_t, mid_t,
       expr('a'), out_t, expr('m'))


@pytest.mark.parametrize('seed', range(24))
def test_random_fixed_point_programs(seed):
  """The native NumPy evaluation, NumPy on the rewritten program and g++ on the
  rewritten program agree bit for bit on seeded random fixed-point programs."""
  st = sodac.compile_source(_random_fixed_program(seed))
  extent = (40, 11)
  rng = np.random.default_rng(seed)
  dtype = golden.np_dtype(st.input_stmts[0].haoda_type)
  inputs = {'a': rng.integers(0, 2**31, extent[::-1]).astype(dtype)}
  native = golden.run(st, inputs)
  low = widths.lower(fixed_point.lower(st))
  assert not fixed_point.has_fixed_types(low)
  rewritten = golden.run(low, inputs)
  compiled = emit_cpp.Oracle(st).run(inputs)
  index = common.box_index(st.valid_box('b', extent))
  assert native['b'][index].size > 0
  assert np.array_equal(native['b'][index], rewritten['b'][index])
  assert np.array_equal(native['b'][index], compiled['b'][index])
