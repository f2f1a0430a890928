"""The DSL's ``half`` element type.

The reference's grammar accepts ``half`` next to ``float`` and ``double``
(src/soda/grammar.py:37-45) and prints it as the HLS ``half`` type, an IEEE
binary16 whose every operation rounds once.  None of the reference's tests/src
programs uses it, so the two programs here live in tests/src_extra.  The rule
set: ``half (op) half`` is a binary16 operation, ``half (op) float / double``
is a float / double operation, an integer operand converts to half - C++'s
usual arithmetic conversions for ``_Float16``.

Three independent evaluators must agree bit for bit: NumPy float16
(oracle/golden.py), g++ ``_Float16`` with one explicit rounding per operation
(oracle/emit_cpp.py), and the CUDA templates with soda::half_t
(soda_b200/csrc/soda_half.cuh) - under CPU emulation here, on the GPU with
``-m gpu``."""
import os

import numpy as np
import pytest

from oracle import emit_cpp, golden
from soda_b200 import ir, sodac
from soda_b200.codegen.cuda import launcher, plan
from tests import common
from tests.emu import build_emu

EXTRA = os.path.join(common.ROOT, 'tests', 'src_extra')
CASES = [('smooth2d_half', (150, 37), {}),
         ('smooth2d_half', (97, 40), {'time_block': 2}),
         ('mix3d_half', (40, 21, 13), {}),
         # binary16 pairs (H2): two cells per HADD2 / HMUL2
         ('jacobi2d_half', (300, 37), {}),
         ('jacobi2d_half', (280, 33), {'time_block': 2}),
         ('jacobi2d_half', (1100, 21), {'time_block': 2,
                                        'options': {'cells': 16}}),
         ('jacobi2d_half', (150, 30), {'time_block': 2,
                                       'options': {'no_pack': True}}),
         # 3-D plans are unpacked by default; pairs on request
         ('jacobi3d_half', (60, 21, 11), {'time_block': 2}),
         ('jacobi3d_half', (60, 19, 12), {'time_block': 2,
                                          'options': {'pack': True}})]


def stencil(name, **overrides):
  with open(os.path.join(EXTRA, name + '.soda')) as fp:
    return sodac.compile_source(fp.read(), **overrides)


def test_type_rules():
  half = ir.Type('half')
  assert half.is_float and half.is_executable and half.width_in_bits == 16
  assert half.numpy_name == 'float16' and half.c_type == 'soda::half_t'
  for other, want in (('int16', 'half'), ('uint64', 'half'), ('half', 'half'),
                      ('float', 'float'), ('double', 'double')):
    assert str(ir.common_type(half, ir.Type(other))) == want
    assert str(ir.common_type(ir.Type(other), half)) == want
  st = stencil('smooth2d_half')
  assert str(sodac.compile_source(str(st))) == str(st)
  # 16-bit cells: eight per lane; a division and a float literal keep the
  # scalar path
  p = plan.make_tuned_pass_plan(st, 2)
  assert p.cells == 8 and p.pack == 1
  assert not plan.packable(st)


def test_half_pairs_are_planned_for_add_sub_mul_programs():
  st = stencil('jacobi2d_half')
  assert plan.packable(st)
  p = plan.make_tuned_pass_plan(st, 2)
  assert p.cells == 8 and p.pack == 2
  # 16-cell lanes on request: the 512-cell strip is two TMA boxes
  wide = plan.make_tuned_pass_plan(st, 2, {'cells': 16})
  assert wide.cells == 16 and wide.strip == 512 and wide.pack == 2
  assert plan.make_tuned_pass_plan(st, 2, {'no_pack': True}).pack == 1
  # `* 0.2f` is a float multiplication of the converted sum: not a half pair
  text = open(os.path.join(EXTRA, 'jacobi2d_half.soda')).read()
  assert not plan.packable(sodac.compile_source(
      text.replace('half(0.2f)', '0.2f')))
  # an integer literal converts to half and is broadcast
  assert plan.packable(sodac.compile_source(text.replace('half(0.2f)', '3')))


def test_every_operation_rounds_once():
  """``a + b + c`` in half is not the rounded float sum: the g++ oracle must
  not be fooled by excess precision."""
  st = sodac.compile_source('''kernel: sum3
burst width: 64
unroll factor: 2
iterate: 1
input half: x(8, *)
output half: y(0, 0) = x(0, 0) + x(1, 0) + x(2, 0)
''')
  # 2048 + 1 + 1: each addition is a tie that rounds back to 2048 in
  # binary16, while the exact (or float) sum 2050 is representable
  x = np.zeros((3, 8), dtype=np.float16)
  x[:, 0], x[:, 1], x[:, 2] = 2048, 1, 1
  for run in (lambda: golden.run(st, {'x': x}),
              lambda: emit_cpp.Oracle(st).run({'x': x})):
    assert float(run()['y'][1, 0]) == 2048.0
  assert float(np.float16(np.float32(2048) + 1 + 1)) == 2050.0


@pytest.mark.parametrize('name,extent,kwargs', CASES)
def test_oracles_agree(name, extent, kwargs):
  st = stencil(name)
  inputs = common.make_inputs(st, extent, seed=11)
  a = golden.run(st, inputs)
  b = emit_cpp.Oracle(st).run(inputs)
  for out in st.output_names:
    index = common.box_index(st.valid_box(out, extent))
    assert a[out].dtype == np.float16
    assert np.array_equal(a[out][index].view(np.uint16),
                          b[out][index].view(np.uint16))
    # the data exercises the type: many distinct finite values
    assert np.isfinite(a[out][index].astype(np.float32)).all()
    assert len(np.unique(a[out][index])) > 500


@pytest.mark.parametrize('name,extent,kwargs', CASES)
def test_under_emulation(name, extent, kwargs):
  st = stencil(name)
  prog = launcher.CudaProgram(build_emu.build_emu_library(st, **kwargs))
  assert str(prog.output_dtypes[0]) == 'float16'
  inputs = common.make_inputs(st, extent, seed=12)
  outputs = {n: np.full(extent[::-1], 77, dtype=np.float16)
             for n in st.output_names}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs,
                               emit_cpp.Oracle(st).run(inputs), sentinel=77)


@pytest.mark.gpu
@pytest.mark.parametrize('name,extent,kwargs', [
    ('smooth2d_half', (1000, 211), {'time_block': 2}),
    ('smooth2d_half', (333, 90), {}),
    ('mix3d_half', (150, 45, 21), {}),
    ('jacobi2d_half', (1000, 300), {'time_block': 2}),
    ('jacobi2d_half', (2000, 150), {'time_block': 2,
                                    'options': {'no_pack': True}}),
    ('jacobi3d_half', (300, 40, 30), {'time_block': 2}),
    ('jacobi3d_half', (300, 40, 30), {'time_block': 2,
                                      'options': {'pack': True}}),
])
def test_on_gpu(name, extent, kwargs):
  from soda_b200.codegen import cuda as cuda_backend
  st = stencil(name)
  prog = cuda_backend.compile_stencil(st, **kwargs)
  inputs = common.make_inputs(st, extent, seed=13)
  outputs = {n: np.full(extent[::-1], 77, dtype=np.float16)
             for n in st.output_names}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs,
                               emit_cpp.Oracle(st).run(inputs), sentinel=77)
