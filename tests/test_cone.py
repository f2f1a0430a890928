"""The dependency-cone checker (oracle/cone.py) against full golden runs: a
window of the full run equals the golden loops run on the window's cone, for
centred, one-sided and multi-input windows; a corrupted cell is reported."""
import numpy as np
import pytest

from oracle import cone
from tests import common


@pytest.mark.parametrize('name,extent,overrides', [
    ('jacobi2d', (200, 150), dict(iterate=5)),
    ('blur', (300, 90), dict(iterate=2)),
    ('denoise2d', (120, 80), {}),
    ('xcorr', (100, 90), {}),
    ('heat3d', (60, 50, 40), dict(iterate=3)),
    ('denoise3d', (50, 40, 30), {}),
])
def test_cone_windows_equal_full_golden_run(name, extent, overrides):
  st = common.stencil(name, **overrides)
  inputs = common.make_inputs(st, extent, seed=2)
  full = common.oracle_outputs(st, inputs)
  size = (32, 24) if st.dim == 2 else (16, 12, 8)
  report = cone.check_host_arrays(st, inputs, full, count=6, size=size, seed=1)
  assert report['bit_exact'] and report['windows'] == 6 * len(st.output_names)
  assert report['cells'] > 0


def test_cone_checker_reports_a_wrong_cell():
  st = common.stencil('jacobi2d', iterate=4)
  extent = (160, 120)
  inputs = common.make_inputs(st, extent, seed=5)
  full = common.oracle_outputs(st, inputs)
  (x0, x1), (y0, y1) = st.valid_box('t0', extent)
  # the windows at the two corners of the valid box are always drawn
  full['t0'][y1 - 1, x1 - 1] = np.nextafter(full['t0'][y1 - 1, x1 - 1],
                                            np.float32(2))
  report = cone.check_host_arrays(st, inputs, full, count=4, size=(16, 16))
  assert not report['bit_exact']
  assert report['first_mismatch']['differing_cells'] == 1


def test_required_windows_straddle_a_seam():
  valid = ((2, 98), (2, 398))
  windows = cone.draw_windows(valid, (16, 16), 3, seed=0,
                              required=[(40, 200 - 8)])
  assert len(windows) == 4
  assert windows[-1] == ((40, 56), (192, 208))
  assert all(valid[d][0] <= w[d][0] < w[d][1] <= valid[d][1]
             for w in windows for d in range(2))
