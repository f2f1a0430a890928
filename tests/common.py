"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

from oracle import emit_cpp, golden
from soda_b200 import sodac

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC_DIR = os.path.join(ROOT, 'tests', 'src')
PROGRAMS = sorted(
    os.path.splitext(os.path.basename(p))[0]
    for p in glob.glob(os.path.join(SRC_DIR, '*.soda')))
PROGRAMS_2D = [
    'blur', 'contrast', 'denoise2d', 'erosion', 'jacobi2d', 'seidel2d',
    'sobel2d', 'xcorr'
]
PROGRAMS_3D = ['denoise3d', 'heat3d', 'jacobi3d']


def source(name: str) -> str:
  with open(os.path.join(SRC_DIR, name + '.soda')) as fp:
    return fp.read()


def stencil(name: str, **overrides):
  return sodac.compile_source(source(name), **overrides)


def make_inputs(st, extent, seed=0, pattern='random'):
  """Seeded inputs.  Integer programs are range-limited so that no C++
  intermediate leaves int32: |x| < 2^10 for the wide
  windows, full uint16 range for blur."""
  if pattern == 'reference':
    return golden.reference_inputs(st, extent, seed)
  rng = np.random.default_rng(seed)
  shape = tuple(extent[::-1])
  result = {}
  for stmt in st.input_stmts:
    dtype = golden.np_dtype(stmt.haoda_type)
    if stmt.haoda_type.is_float:
      result[stmt.name] = rng.random(shape, dtype=np.float32).astype(dtype)
    elif st.app_name == 'blur':
      result[stmt.name] = rng.integers(0, 65536, shape).astype(dtype)
    elif stmt.haoda_type.is_signed:
      result[stmt.name] = rng.integers(-1024, 1024, shape).astype(dtype)
    else:
      result[stmt.name] = rng.integers(0, 1024, shape).astype(dtype)
  return result


def box_index(box):
  return tuple(slice(lo, hi) for lo, hi in reversed(box))


def assert_matches_oracle(st, extent, got, want, sentinel=None):
  """Bit-exact inside each output's valid box; untouched outside."""
  for name in st.output_names:
    box = st.valid_box(name, extent)
    if any(hi <= lo for lo, hi in box):
      inside = None
    else:
      inside = box_index(box)
      a = np.ascontiguousarray(got[name][inside])
      b = np.ascontiguousarray(want[name][inside])
      if not np.array_equal(a.view(np.uint8), b.view(np.uint8)):
        bad = np.argwhere(a != b)
        raise AssertionError(
            '%s: %d of %d cells differ, first at %s: got %r want %r' %
            (name, len(bad), a.size, bad[0].tolist(), a[tuple(bad[0])],
             b[tuple(bad[0])]))
    if sentinel is not None:
      mask = np.ones(got[name].shape, dtype=bool)
      if inside is not None:
        mask[inside] = False
      assert np.all(got[name][mask] == sentinel), \
          '%s: cells outside the valid box were written' % name


def make_params(st, seed=0, pattern='random'):
  """``param`` arrays: the reference test main's ``p[x][y] = x + y`` or seeded
  random values."""
  if pattern == 'reference':
    return golden.reference_params(st)
  rng = np.random.default_rng(seed + 1000)
  result = {}
  for stmt in st.param_stmts:
    shape = tuple(int(x) for x in stmt.size)
    dtype = golden.np_dtype(stmt.haoda_type)
    if stmt.haoda_type.is_float:
      result[stmt.name] = (rng.random(shape) - 0.5).astype(dtype)
    else:
      result[stmt.name] = rng.integers(-4, 5, shape).astype(dtype)
  return result


def oracle_outputs(st, inputs, use_cpp=True, params=None):
  if use_cpp:
    return emit_cpp.Oracle(st).run(inputs, params=params)
  return golden.run(st, inputs, params=params)
