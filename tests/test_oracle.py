"""The oracle against itself and against known answers.

Three independent restatements of the reference's golden loops exist:
oracle/golden.py (NumPy), oracle/emit_cpp.py (g++) and the per-cell Python loop
below.  The g++ one is additionally judged by the reference's own generated
test main (oracle/pin_against_reference.py -> tests/golden/reference_pin.log).
"""
import os

import numpy as np
import pytest

from oracle import emit_cpp, golden
from tests import common


@pytest.mark.parametrize('name', common.PROGRAMS)
def test_numpy_and_cpp_oracles_agree(name):
  st = common.stencil(name)
  extent = list(golden.default_extent(st))
  extent[-1] += 5
  for pattern in ('reference', 'random'):
    inputs = common.make_inputs(st, extent, seed=2, pattern=pattern)
    a = golden.run(st, inputs)
    b = emit_cpp.Oracle(st).run(inputs)
    common.assert_matches_oracle(st, extent, a, b)


def test_reference_pin_log_is_green():
  """Committed verdicts of the reference's print_test judging the oracle."""
  path = os.path.join(common.ROOT, 'tests', 'golden', 'reference_pin.log')
  lines = [l for l in open(path).read().splitlines() if not l.startswith('#')]
  assert len(lines) >= 23
  assert all(line.startswith('PASS') for line in lines), lines


def test_blur_known_answer():
  """input = p + q  =>  blur_x = p + q + 1, blur_y = p + q + 2 (hand-derived:
  the mean of three consecutive integers is the middle one)."""
  st = common.stencil('blur')
  extent = (50, 9)
  inputs = golden.reference_inputs(st, extent)
  q, p = np.indices(extent[::-1])
  out = emit_cpp.Oracle(st).run(inputs)['blur_y']
  (x0, x1), (y0, y1) = st.valid_box('blur_y', extent)
  assert (x0, x1, y0, y1) == (0, 48, 0, 7)
  assert np.array_equal(out[y0:y1, x0:x1], (p + q + 2)[y0:y1, x0:x1])
  assert np.all(out[y1:, :] == 0) and np.all(out[:, x1:] == 0)


def test_blur_iterate2_known_answer():
  st = common.stencil('blur', iterate=2)
  extent = (50, 9)
  q, p = np.indices(extent[::-1])
  out = golden.run(st, golden.reference_inputs(st, extent))['blur_y']
  (x0, x1), (y0, y1) = st.valid_box('blur_y', extent)
  assert (x0, x1, y0, y1) == (0, 46, 0, 5)
  assert np.array_equal(out[y0:y1, x0:x1], (p + q + 4)[y0:y1, x0:x1])


def test_blur_truncating_division():
  """C++ integer division truncates; uint16 operands are promoted to int so
  3 * 65535 does not wrap before the division."""
  st = common.stencil('blur')
  grid = np.full((5, 40), 65535, dtype=np.uint16)
  grid[:, ::2] = 65534
  out = golden.run(st, {'input': grid})['blur_y']
  # blur_x = (65534|65535 thrice) / 3 exactly; blur_y = (a + b + a) / 3 truncated
  assert out[0, 0] == (65534 + 65535 + 65534) // 3
  assert out[0, 1] == (65535 + 65534 + 65535) // 3


def test_erosion_and_xcorr_known_answers():
  """input = p + q: the 19-wide min picks the lowest corner, and the 19x19 box
  sum is 361 (p + q) before the int16 store wraps it."""
  extent = (60, 25)
  q, p = np.indices(extent[::-1]).astype(np.int64)
  st = common.stencil('erosion')
  out = emit_cpp.Oracle(st).run(golden.reference_inputs(st, extent))['output']
  inside = common.box_index(st.valid_box('output', extent))
  assert np.array_equal(out[inside], (p + q - 18)[inside])
  st = common.stencil('xcorr')
  out = emit_cpp.Oracle(st).run(golden.reference_inputs(st, extent))['tmp3']
  inside = common.box_index(st.valid_box('tmp3', extent))
  wrap16 = lambda v: ((v + 2**15) % 2**16) - 2**15
  tmp2 = wrap16(wrap16(19 * (p + q)) * 1 * 19)  # both int16 stores wrap
  value = (tmp2 - (p + q)) * (p + q)
  want = wrap16(np.sign(value) * (np.abs(value) // 256))  # C++ truncation
  assert np.array_equal(out[inside], want[inside])


def test_jacobi_fixed_point_on_linear_field():
  """A linear integer-valued field is a fixed point of the 5-point average:
  the sum is exactly 5v and fl(5v * 0.2f) == v for small integers."""
  st = common.stencil('jacobi2d', iterate=3)
  extent = (40, 30)
  q, p = np.indices(extent[::-1])
  field = (3 * p + 2 * q).astype(np.float32)
  out = golden.run(st, {'t1': field})['t0']
  inside = common.box_index(st.valid_box('t0', extent))
  assert st.valid_box('t0', extent) == ((3, 37), (3, 27))
  assert np.array_equal(out[inside], field[inside])


def test_sqrt_is_evaluated_in_double():
  """g++ resolves the generated `sqrt(x)` (only <cmath> included) to
  ::sqrt(double): decltype(sqrt(1.0f)) is double with g++ 13.3.  Both oracles
  and the CUDA functors follow that."""
  st = common.stencil('denoise2d')
  extent = (32, 8)
  inputs = common.make_inputs(st, extent, seed=5)
  full = golden.run(st, inputs, keep_intermediates=True)
  u = inputs['u'].astype(np.float32)
  g = full['g']
  y, x = 3, 7
  du, dd = u[y, x] - u[y - 1, x], u[y, x] - u[y + 1, x]
  dl, dr = u[y, x] - u[y, x - 1], u[y, x] - u[y, x + 1]
  s = np.float32(1.0) + du * du + dd * dd + dl * dl + dr * dr   # float32 chain
  want = np.float32(np.float64(np.float32(1.0)) / np.sqrt(np.float64(s)))
  assert g[y, x] == want


def test_valid_box_and_untouched_border():
  st = common.stencil('jacobi2d')
  assert st.valid_box('t0', (32, 6)) == ((2, 30), (2, 4))
  assert st.valid_box('t1_iter1', (32, 6)) == ((1, 31), (1, 5))
  out = golden.run(st, common.make_inputs(st, (32, 6)))['t0']
  assert np.count_nonzero(out[:2]) == 0 and np.count_nonzero(out[4:]) == 0


def test_default_extents_match_reference_main():
  """sizes of the reference's default test run."""
  want = {'blur': (2000, 4), 'contrast': (480, 18), 'denoise2d': (32, 6),
          'denoise3d': (32, 32, 6), 'erosion': (480, 20),
          'heat3d': (32, 32, 6), 'jacobi2d': (32, 6), 'jacobi3d': (32, 32, 6),
          'seidel2d': (32, 6), 'sobel2d': (32, 4), 'xcorr': (480, 20)}
  for name, extent in want.items():
    assert golden.default_extent(common.stencil(name)) == extent


def test_reference_compare_criterion():
  """src/soda/codegen/frt/host.py:633-657: ints exact; floats fail only when
  both the absolute and the relative error exceed 1e-5."""
  from oracle import compare
  want = np.array([1.0, 1000.0, 1e-7, 0.0], dtype=np.float32)
  got = want + np.array([5e-6, 5e-3, 5e-6, 2e-5], dtype=np.float32)
  # 5e-6 abs ok; 5e-3 on 1000 is 5e-6 relative ok; 5e-6 abs ok; 2e-5 on 0: fail
  assert compare.error_count(got, want) == 1
  assert compare.error_count(got, want, threshold=0) == 4
  ints = np.array([1, 2, 3], dtype=np.int16)
  assert compare.error_count(ints, ints) == 0
  assert compare.error_count(ints, ints + np.int16(1)) == 3
  a = np.float32(1.0)
  b = np.nextafter(a, np.float32(2.0))
  assert compare.ulp_distance(np.array([a]), np.array([b]))[0] == 1


@pytest.mark.parametrize('name', ['denoise2d', 'denoise3d'])
def test_math_precision_modes(name):
  """``sqrt(float)``: through double (default: what g++ makes of the reference's
  generated code with plain <cmath>) or in float (``--math-precision float``:
  the std:: float overloads in scope, as Xilinx's headers may arrange).  Both
  evaluators implement both readings and agree bit for bit in each; the two
  readings differ from each other (so the switch is not vacuous)."""
  import numpy as np
  from oracle import emit_cpp, golden
  from soda_b200 import ir, sodac
  results = {}
  for mode in ('double', 'float'):
    st = sodac.compile_source(common.source(name), math_precision=mode)
    g = [s for s in st.local_stmts if s.name == 'g'][0]
    inner = ir.unwrap(g.expr) if mode == 'float' else g.expr
    assert (g.expr.haoda_type == ir.FLOAT) and \
        (isinstance(g.expr, ir.Cast) == (mode == 'double')), inner
    extent = list(golden.default_extent(st))
    extent[0] += 9
    extent[-1] += 7
    inputs = common.make_inputs(st, extent, seed=17)
    a = golden.run(st, inputs)['output']
    b = emit_cpp.Oracle(st).run(inputs)['output']
    inside = common.box_index(st.valid_box('output', extent))
    assert np.array_equal(a[inside].view(np.uint32), b[inside].view(np.uint32))
    results[mode] = a[inside]
  assert not np.array_equal(results['double'].view(np.uint32),
                            results['float'].view(np.uint32))
  assert np.allclose(results['double'], results['float'], rtol=1e-5)
