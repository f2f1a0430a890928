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


@pytest.mark.parametrize('name,kwargs', [
    ('jacobi2d', dict(extent=(70, 200), time_block=2, iterate=5)),
    ('heat3d', dict(extent=(40, 12, 90), time_block=2, iterate=4,
                    options={'rows': 8})),
    ('blur', dict(extent=(80, 33), iterate=2)),  # chunks of a few rows
])
@pytest.mark.parametrize('host_chunks', [-8, -3])
def test_shrinking_host_chunks_under_emulation(name, kwargs, host_chunks):
  """A negative chunk count selects the layout the pipeline chooses by itself
  for large grids: every chunk 8 % shorter than the one before it (never
  shorter than the reach of all passes while the grid allows), upload pieces
  that end where the chunk windows end."""
  run_case(name, host_chunks=host_chunks, **kwargs)


POW2_2D = '''kernel: pow2_2d
burst width: 64
unroll factor: 2
iterate: 3
input float: a(32, *)
output float: b(0, 0) = .5f * a(0, 0) + a(1, 0) * .125f + .3f * a(-1, 0) - .125f * a(0, 1) + (a(0, -1) * .125f)
'''


@pytest.mark.parametrize('name,kwargs', [
    ('heat3d', dict(extent=(40, 12, 20), time_block=2, iterate=4,
                    options={'rows': 8, 'pow2_fma': True})),
    (POW2_2D, dict(extent=(300, 40), time_block=3,
                   options={'pow2_fma': True})),
    (POW2_2D, dict(extent=(300, 40), time_block=1,
                   options={'pow2_fma': True, 'no_pack': True})),
])
def test_fused_power_of_two_coefficients_under_emulation(name, kwargs):
  """--cuda-pow2-fma (soda::fma_pow2): fused `c * x + acc` for power-of-two
  literals only (`.3f * a` stays a product and a sum); equal to the un-fused
  oracle bit for bit on data in the normal range, scalar and packed."""
  from soda_b200 import sodac
  from soda_b200.codegen.cuda import emit
  if name == POW2_2D:
    st = sodac.compile_source(name)
    source = emit.emit_program(st, kwargs['time_block'], kwargs['options'])
    assert source.count('soda::fma_pow2(') == 3  # the first term stays a product
    assert '(.3f * ' in source
    prog = launcher.CudaProgram(build_emu.build_emu_library(
        st, time_block=kwargs['time_block'], options=kwargs['options']))
    extent = kwargs['extent']
    inputs = common.make_inputs(st, extent, seed=5)
    outputs = {n: np.full(extent[::-1], 77, dtype=d)
               for n, d in zip(prog.output_names, prog.output_dtypes)}
    prog.run_host(inputs, outputs)
    common.assert_matches_oracle(st, extent, outputs,
                                 common.oracle_outputs(st, inputs), sentinel=77)
  else:
    run_case(name, **kwargs)


ONE_SIDED_2D_NEG = '''kernel: one_sided_neg
burst width: 64
unroll factor: 2
iterate: 2
input float: a(32, *)
output float: b(0, 0) = (a(-1, -2) + a(-2, -1)) * 0.5f
'''
ONE_SIDED_2D_POS = '''kernel: one_sided_pos
burst width: 64
unroll factor: 2
iterate: 3
input int32: a(32, *)
output int32: b(0, 0) = (a(1, 2) + a(2, 1)) / 2
'''
ONE_SIDED_3D = '''kernel: one_sided_3d
burst width: 64
unroll factor: 2
iterate: 2
input float: a(32, 32, *)
output float: b(0, 0, 0) = (a(0, 1, 2) + a(1, 0, 1)) * 0.5f
'''


@pytest.mark.parametrize('text,extent', [(ONE_SIDED_2D_NEG, (70, 36)),
                                         (ONE_SIDED_2D_POS, (70, 36)),
                                         (ONE_SIDED_3D, (40, 12, 20))])
@pytest.mark.parametrize('host_chunks', [1, 3])
def test_one_sided_windows_over_several_passes(text, extent, host_chunks):
  """A window that lies strictly on one side of the stored cell in the
  streamed dimension: the pass reach the runtime sums over the passes (chunk
  windows, halo depths) must never go negative.  iterate > time_block, so
  intermediates travel through the scratch arrays."""
  from soda_b200 import sodac
  st = sodac.compile_source(text)
  prog = launcher.CudaProgram(
      build_emu.build_emu_library(st, time_block=1, options={'rows': 8}
                                  if st.dim == 3 else None))
  for index in range(prog.num_passes):
    info = prog.pass_info(index)
    assert all(lo <= 0 <= hi for lo, hi in zip(info.reach_lo, info.reach_hi))
  inputs = common.make_inputs(st, extent, seed=3)
  outputs = {n: np.full(extent[::-1], 77, dtype=d)
             for n, d in zip(prog.output_names, prog.output_dtypes)}
  prog.run_host(inputs, outputs,
                opts=launcher.make_opts(host_chunks=host_chunks))
  common.assert_matches_oracle(st, extent, outputs,
                               common.oracle_outputs(st, inputs), sentinel=77)


@pytest.mark.parametrize('name,kwargs', [
    ('denoise2d', dict(extent=(100, 30))),
    ('denoise3d', dict(extent=(60, 19, 7), options={'rows': 8})),
])
def test_float_math_mode_under_emulation(name, kwargs):
  """--math-precision float: sqrtf in the functors, std::sqrt(float) in the
  oracle; bit-exact through the templates like the default mode."""
  run_case(name, math_precision='float', **kwargs)


def test_several_warp_ctas_on_a_small_grid():
  """Small grids run as one-warp CTAs (soda_runtime.cuh, narrow_grid_strips);
  SODA_CUDA_NARROW_STRIPS=0 keeps the program's own CTA shape, so the
  several-warp path is exercised under emulation too.  Run in a fresh process:
  the threshold is read once."""
  import os
  import subprocess
  import sys
  code = (
      "import numpy as np\n"
      "from tests import test_emu_kernels as t\n"
      "t.run_case('jacobi2d', extent=(1100, 30), time_block=2, iterate=5,"
      " segment=12)\n"
      "t.run_case('blur', extent=(1300, 21), time_block=2, iterate=2)\n"
      "print('ok')\n")
  env = dict(os.environ, SODA_CUDA_NARROW_STRIPS='0')
  result = subprocess.run([sys.executable, '-c', code], cwd=common.ROOT,
                          env=env, capture_output=True, text=True, timeout=900)
  assert result.returncode == 0 and 'ok' in result.stdout, result.stderr[-2000:]


def test_one_dimensional_program_under_emulation():
  """A 1-D program runs lifted to an N x 1 grid (optimization/lift.py): same
  values and valid range as the 1-D golden loops, through the program-named
  entry point with 1-D quadruples and through the generic 2-D one."""
  import os
  from soda_b200 import sodac
  with open(os.path.join(common.ROOT, 'tests', 'src_extra',
                         'smooth1d.soda')) as fp:
    st = sodac.compile_source(fp.read())
  assert st.dim == 1
  prog = launcher.CudaProgram(build_emu.build_emu_library(st, time_block=2))
  assert (prog.dim, prog.source_dim, prog.num_passes) == (2, 1, 2)
  for n in (700, 5, 7):
    x = np.random.default_rng(n).random((n,), dtype=np.float32)
    want = common.oracle_outputs(st, {'a': x})['b']
    (lo, hi), = st.valid_box('b', (n,))
    for app_entry in (True, False):
      out = np.full((n,), 77, dtype=np.float32)
      prog.run_host({'a': x}, {'b': out}, use_app_entry=app_entry)
      if hi > lo:
        assert np.array_equal(out[lo:hi].view(np.uint32),
                              want[lo:hi].view(np.uint32))
      assert np.all(out[:max(lo, 0)] == 77) and np.all(out[max(hi, lo):] == 77)
