"""ORACLE (test infrastructure): the reference's own pass/fail criterion.

The generated test main of the reference compares the kernel's output with its
golden loops element by element (reference:
src/soda/codegen/frt/host.py:625-667): integers must be equal; a float fails
only if BOTH its absolute error and its relative error exceed the threshold
(1e-5 by default, ``THRESHOLD`` in the environment overrides it; both are
compared squared, as there).
"""
import os

import numpy as np


def error_count(got: np.ndarray, want: np.ndarray, threshold=None) -> int:
  """Number of elements the reference's test main would report."""
  if got.shape != want.shape:
    raise ValueError('shape mismatch')
  if got.dtype.kind != 'f':
    return int(np.count_nonzero(got != want))
  if threshold is None:
    threshold = float(os.environ.get('THRESHOLD', '0.00001'))
  threshold *= threshold
  fpga = got.astype(np.float64)
  cpu = want.astype(np.float64)
  diff2 = (fpga - cpu) * (fpga - cpu)
  with np.errstate(divide='ignore', invalid='ignore'):
    relative = diff2 / (cpu * cpu)
  return int(np.count_nonzero((diff2 > threshold) & (relative > threshold)))


def ulp_distance(got: np.ndarray, want: np.ndarray) -> np.ndarray:
  """Distance in units in the last place between two float32 arrays."""
  def ordered(a):
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.where(bits < 0, -(bits & 0x7fffffff), bits)
  return np.abs(ordered(got) - ordered(want))
