"""The real kernel templates, plan tables and C-ABI runtime executed on the CPU
by the test-only emulation in tests/emu (every CUDA thread is a std::thread;
see soda_emu.h).  This is how index logic is checked in a container without a
GPU; the `-m gpu` tests repeat the comparison on a B200."""
import numpy as np
import pytest

from oracle import golden
from soda_b200.codegen.cuda import launcher
from tests import common
from tests.emu import build_emu


def run_case(name, extent=None, time_block=None, options=None, segment=0,
             seed=1, host_chunks=0, **overrides):
  st = common.stencil(name, **overrides)
  prog = launcher.CudaProgram(
      build_emu.build_emu_library(st, time_block=time_block, options=options))
  extent = tuple(extent or golden.default_extent(st))
  inputs = common.make_inputs(st, extent, seed=seed)
  outputs = {
      n: np.full(extent[::-1], 77, dtype=d)
      for n, d in zip(prog.output_names, prog.output_dtypes)
  }
  prog.run_host(inputs, outputs,
                opts=launcher.make_opts(segment=segment,
                                        host_chunks=host_chunks))
  common.assert_matches_oracle(st, extent, outputs,
                               common.oracle_outputs(st, inputs), sentinel=77)


CASES_2D = [
    ('jacobi2d', dict(extent=(300, 70), time_block=2, iterate=5, segment=24)),
    ('jacobi2d', dict(extent=(260, 50), time_block=4, iterate=9, segment=13)),
    ('blur', dict(extent=(700, 33), time_block=2, iterate=2, segment=10)),
    ('sobel2d', dict(extent=(257, 21), segment=7)),
    ('denoise2d', dict(extent=(150, 40), segment=16)),
    ('seidel2d', dict(extent=(131, 37), time_block=2, iterate=4, segment=9)),
    ('xcorr', dict()),
    ('jacobi2d', dict(extent=(3, 3))),
]
CASES_2D.append(('jacobi2d', dict(extent=(300, 40), time_block=3, iterate=3,
                                  options={'no_pack': True})))
# 16-cell lanes of 16-bit cells: the 512-cell strip is two TMA boxes side by
# side (three strips at this width, the last one ragged)
CASES_2D.append(('blur', dict(extent=(1300, 21), time_block=2, iterate=2,
                              options={'cells': 16}, segment=8)))
CASES_3D = [
    ('jacobi3d', dict(extent=(150, 21, 11), time_block=2, iterate=3,
                      options={'rows': 8}, segment=6)),
    ('heat3d', dict(extent=(40, 30, 9), time_block=1, iterate=2,
                    options={'rows': 4}, segment=5)),
    ('denoise3d', dict(extent=(140, 19, 7), options={'rows': 8}, segment=4)),
    # patches: every thread owns cy rows of the tile
    ('jacobi3d', dict(extent=(150, 25, 11), time_block=2, iterate=3,
                      options={'rows': 16, 'cy': 4}, segment=6)),
    ('heat3d', dict(extent=(40, 30, 9), time_block=1, iterate=2,
                    options={'rows': 8, 'cy': 2}, segment=5)),
    ('heat3d', dict(extent=(40, 30, 9), time_block=3, iterate=3,
                    options={'rows': 16, 'cy': 2, 'no_pack': True})),
    ('denoise3d', dict(extent=(140, 19, 7), options={'rows': 8, 'cy': 2},
                       segment=4)),
    ('denoise3d', dict(extent=(40, 33, 6), options={'rows': 16, 'cy': 4})),
]


@pytest.mark.parametrize('name,kwargs', CASES_2D)
def test_2d_templates_under_emulation(name, kwargs):
  run_case(name, **kwargs)


@pytest.mark.parametrize('name,kwargs', CASES_3D)
def test_3d_templates_under_emulation(name, kwargs):
  run_case(name, **kwargs)


@pytest.mark.parametrize('name,kwargs', [
    ('jacobi2d', dict(extent=(150, 40), time_block=2, iterate=4, segment=16)),
    ('seidel2d', dict(extent=(140, 30), time_block=2, iterate=2)),
    ('heat3d', dict(extent=(40, 20, 12), time_block=2, iterate=2,
                    options={'rows': 8})),
])
def test_computation_reuse_stages_under_emulation(name, kwargs):
  """--computation-reuse adds cr_var stages; they are ordinary nodes of the
  fused DAG (shared partial sums in registers / shuffles)."""
  run_case(name, computation_reuse='yes', **kwargs)


@pytest.mark.parametrize('name,kwargs', [
    ('jacobi2d', dict(extent=(150, 60), time_block=2, iterate=5, segment=16)),
    ('denoise2d', dict(extent=(64, 40))),
    ('heat3d', dict(extent=(40, 12, 20), time_block=2, iterate=4,
                    options={'rows': 8})),
    ('jacobi2d', dict(extent=(40, 7), time_block=2)),  # chunks thinner than reach
])
def test_pipelined_host_path_under_emulation(name, kwargs):
  """soda_cuda_plan_run_host cuts the grid into overlapping chunks along the
  streamed dimension so that copies overlap compute; results must not change."""
  run_case(name, host_chunks=3, **kwargs)
