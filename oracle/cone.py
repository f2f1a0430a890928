"""ORACLE (test infrastructure, not product): parity on grids too large for a
full golden run, by dependency cones.

An output cell depends on the input cells inside its overall stencil window
(reference: src/soda/core.py:877-919; all ``iterate`` iterations chained,
src/soda/core.py:320-336).  So a window of the output can be checked exactly by
running the golden loops (oracle/emit_cpp.py, the restatement of
src/soda/codegen/frt/host.py:556-624) on the sub-array ``window + overall
stencil window`` of the inputs: the valid box of that small run is exactly the
window, and every value in it went through the same operations in the same
order as in a golden run of the whole grid.

Windows are drawn seeded: a fixed share hugs the corners / edges of the valid
box (where tiles, strips, segments and slabs start and end), the rest is
uniform.  ``required`` lets the caller add windows that straddle seams it knows
about (slab boundaries of a multi-GPU run, chunk seams of the host pipeline).

Only tests/, __graft_entry__.smoke() and bench.py's checking legs may import
this.
"""
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from oracle import emit_cpp

Box = Tuple[Tuple[int, int], ...]  # (lo, hi) per dimension, dimension 0 first


def draw_windows(valid: Box, size: Sequence[int], count: int, seed: int = 0,
                 required: Sequence[Sequence[int]] = ()) -> List[Box]:
  """``count`` windows of ``size`` cells (clipped to the valid box) plus one at
  every ``required`` origin."""
  rng = np.random.default_rng(seed)
  dim = len(valid)
  size = [min(size[d], valid[d][1] - valid[d][0]) for d in range(dim)]
  if any(s <= 0 for s in size):
    return []

  def at(origin) -> Box:
    box = []
    for d in range(dim):
      lo = min(max(origin[d], valid[d][0]), valid[d][1] - size[d])
      box.append((lo, lo + size[d]))
    return tuple(box)

  windows = []
  # the two opposite corners of the valid box, then random origins
  windows.append(at([valid[d][0] for d in range(dim)]))
  windows.append(at([valid[d][1] for d in range(dim)]))
  while len(windows) < count:
    windows.append(at([int(rng.integers(valid[d][0], valid[d][1] - size[d] + 1))
                       for d in range(dim)]))
  windows = windows[:max(count, 0)]
  for origin in required:
    windows.append(at(list(origin)))
  return windows


def check_windows(stencil, extent: Sequence[int],
                  read_input: Callable[[str, Box], np.ndarray],
                  read_output: Callable[[str, Box], np.ndarray],
                  windows: Sequence[Box],
                  params: Optional[Dict[str, np.ndarray]] = None,
                  oracle: Optional[emit_cpp.Oracle] = None) -> Dict:
  """Compares every window of every output bit for bit with the golden loops
  run on the window's dependency cone.

  ``read_input(name, box)`` / ``read_output(name, box)`` return the cells of
  ``box`` (numpy, shape ``box`` sizes reversed: dimension 0 last) - they may
  slice a host array or copy a box back from a device.  Returns
  ``{'windows': n, 'cells': c, 'bit_exact': bool, 'first_mismatch': ...}``.
  """
  oracle = oracle or emit_cpp.Oracle(stencil)
  dim = stencil.dim
  report = {'windows': 0, 'cells': 0, 'bit_exact': True, 'first_mismatch': None}
  for name in stencil.output_names:
    lo, hi = stencil.window_bounds[name]
    valid = stencil.valid_box(name, extent)
    for window in windows:
      window = tuple((max(window[d][0], valid[d][0]), min(window[d][1],
                                                          valid[d][1]))
                     for d in range(dim))
      if any(b <= a for a, b in window):
        continue
      # the cone: cells the window reads, all of which exist in the grid
      # because the window lies inside the valid box
      cone = tuple((window[d][0] + min(lo[d], 0), window[d][1] + max(hi[d], 0))
                   for d in range(dim))
      inputs = {n: np.ascontiguousarray(read_input(n, cone))
                for n in stencil.input_names}
      want = oracle.run(inputs, params=params)[name]
      sub_extent = tuple(b - a for a, b in cone)
      sub_valid = stencil.valid_box(name, sub_extent)
      index = tuple(slice(a, b) for a, b in reversed(sub_valid))
      want = np.ascontiguousarray(want[index])
      got = np.ascontiguousarray(read_output(name, window))
      assert got.shape == want.shape, (got.shape, want.shape, window, cone)
      report['windows'] += 1
      report['cells'] += int(got.size)
      if not np.array_equal(got.view(np.uint8), want.view(np.uint8)):
        report['bit_exact'] = False
        if report['first_mismatch'] is None:
          bad = np.argwhere(got != want)
          where = bad[0].tolist() if len(bad) else []
          report['first_mismatch'] = {
              'output': name,
              'window': [list(w) for w in window],
              'cell': where,
              'got': repr(got[tuple(where)]) if where else None,
              'want': repr(want[tuple(where)]) if where else None,
              'differing_cells': int(len(bad)),
          }
  return report


def host_reader(arrays: Dict[str, np.ndarray]):
  """``read_*`` callback over full-grid numpy arrays."""

  def read(name: str, box: Box) -> np.ndarray:
    return arrays[name][tuple(slice(a, b) for a, b in reversed(box))]

  return read


def check_host_arrays(stencil, inputs: Dict[str, np.ndarray],
                      outputs: Dict[str, np.ndarray], count: int = 16,
                      size: Optional[Sequence[int]] = None, seed: int = 0,
                      required: Sequence[Sequence[int]] = (),
                      params=None) -> Dict:
  """Convenience wrapper for full-grid host arrays."""
  first = inputs[stencil.input_names[0]]
  extent = tuple(first.shape[::-1])
  size = size or ((64, 64) if stencil.dim == 2 else (24, 16, 16))
  # windows are drawn inside the first output's valid box; check_windows clips
  # them to every other output's
  windows = draw_windows(stencil.valid_box(stencil.output_names[0], extent),
                         size, count, seed, required)
  return check_windows(stencil, extent, host_reader(inputs),
                       host_reader(outputs), windows, params)
