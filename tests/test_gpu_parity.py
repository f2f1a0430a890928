"""GPU parity: the CUDA path, called through the C ABI, against the oracle.

Bit-exact for every program (float programs are compiled with --fmad=false, so
they perform exactly the IEEE operations g++ performs for the reference's
generated code)."""
import numpy as np
import pytest

from oracle import golden
from soda_b200.codegen import cuda as cuda_backend
from soda_b200.codegen.cuda import launcher
from tests import common

pytestmark = pytest.mark.gpu

SENTINELS = {'f': 77.0, 'i': 77, 'u': 77}


def run_case(name, extent=None, seed=0, pattern='random', time_block=None,
             options=None, segment=0, host_chunks=0, **overrides):
  st = common.stencil(name, **overrides)
  prog = cuda_backend.compile_stencil(st, time_block=time_block,
                                      options=options)
  extent = tuple(extent or golden.default_extent(st))
  inputs = common.make_inputs(st, extent, seed=seed, pattern=pattern)
  outputs = {
      n: np.full(extent[::-1], 77, dtype=d)
      for n, d in zip(prog.output_names, prog.output_dtypes)
  }
  before = prog.launch_count()
  prog.run_host(inputs, outputs,
                opts=launcher.make_opts(segment=segment,
                                        host_chunks=host_chunks))
  launches = prog.launch_count() - before
  chunks = max(1, host_chunks)
  assert (prog.num_passes - 1) * chunks <= launches <= prog.num_passes * chunks
  want = common.oracle_outputs(st, inputs)
  common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)
  return prog


@pytest.mark.parametrize('name', common.PROGRAMS)
def test_reference_default_size(name):
  """Every tests/src program at the size the reference's test main uses, with
  the reference's integer init pattern p+q(+r)."""
  run_case(name, pattern='reference')


@pytest.mark.parametrize('name', common.PROGRAMS_2D)
def test_2d_many_strips_and_segments(name):
  run_case(name, extent=(1000, 301), seed=3, segment=64)


@pytest.mark.parametrize('name', common.PROGRAMS_3D)
def test_3d_many_tiles_and_segments(name):
  run_case(name, extent=(250, 37, 41), seed=4, segment=16)


@pytest.mark.parametrize('time_block', [1, 2, 3, 4, 8])
def test_jacobi2d_time_blocks(time_block):
  run_case('jacobi2d', extent=(777, 500), seed=5, iterate=11,
           time_block=time_block)


@pytest.mark.parametrize('time_block', [1, 2, 3])
def test_heat3d_time_blocks(time_block):
  run_case('heat3d', extent=(200, 50, 60), seed=6, iterate=5,
           time_block=time_block)


def test_blur_iterate2_config():
  """BASELINE config C1: uint16 blur, tile 2000, iterate 2, bit-exact."""
  for height in (4, 5, 2048):
    run_case('blur', extent=(2000, height), seed=0, iterate=2)
  run_case('blur', extent=(2000, 64), pattern='reference', iterate=2)


def test_empty_valid_box():
  run_case('jacobi2d', extent=(3, 3))
  run_case('jacobi2d', extent=(4, 4))
  run_case('heat3d', extent=(4, 9, 9))


def test_unaligned_extents():
  run_case('jacobi2d', extent=(1, 40))
  run_case('jacobi2d', extent=(129, 5))
  run_case('seidel2d', extent=(131, 37), iterate=4, time_block=2)
  run_case('blur', extent=(1999, 33))
  run_case('sobel2d', extent=(33, 1000))


def test_strided_host_arrays():
  st = common.stencil('jacobi2d')
  prog = cuda_backend.compile_stencil(st)
  extent = (300, 100)
  inputs = common.make_inputs(st, extent, seed=9)
  padded_in = np.zeros((100, 333), dtype=np.float32)
  padded_in[:, :300] = inputs['t1']
  padded_out = np.full((100, 411), 77, dtype=np.float32)
  prog.run_host({'t1': padded_in[:, :300]}, {'t0': padded_out[:, :300]})
  want = common.oracle_outputs(st, inputs)
  common.assert_matches_oracle(st, extent, {'t0': padded_out[:, :300]}, want,
                               sentinel=77)
  assert np.all(padded_out[:, 300:] == 77)


def test_generic_entry_point_and_plan_reuse():
  st = common.stencil('denoise2d')
  prog = cuda_backend.compile_stencil(st)
  extent = (500, 200)
  want = None
  with prog.create_plan(extent) as plan:
    for seed in (1, 2):
      inputs = common.make_inputs(st, extent, seed=seed)
      outputs = {'output': np.full(extent[::-1], 77, dtype=np.float32)}
      plan.run_host(inputs, outputs)
      want = common.oracle_outputs(st, inputs)
      common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)
  outputs = {'output': np.full(extent[::-1], 77, dtype=np.float32)}
  prog.run_host(inputs, outputs, use_app_entry=False)
  common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)


def test_errors_are_reported():
  st = common.stencil('jacobi2d')
  prog = cuda_backend.compile_stencil(st)
  with pytest.raises(launcher.SodaCudaError) as err:
    prog.create_plan((0, 5))
  assert 'extent' in str(err.value)
  bad = np.zeros((10, 10), dtype=np.float32)[:, ::2]
  with pytest.raises(launcher.SodaCudaError):
    prog.run_host({'t1': bad})


def test_jacobi2d_large_linearity_property():
  """Size-independent check at a grid too big for a quick CPU golden run:
  the 5-point average is linear, so run(a) + run(b) == run(a + b) up to float
  rounding, and a constant field is a fixed point (exactly: 5c*0.2f rounds
  back to c for c = 1)."""
  st = common.stencil('jacobi2d', iterate=8)
  prog = cuda_backend.compile_stencil(st, time_block=4)
  extent = (4096, 4096)
  ones = np.ones(extent[::-1], dtype=np.float32)
  out = prog.run_host({'t1': ones})['t0']
  inside = common.box_index(st.valid_box('t0', extent))
  assert np.all(out[inside] == 1.0)
  rng = np.random.default_rng(0)
  a = rng.random(extent[::-1], dtype=np.float32)
  b = rng.random(extent[::-1], dtype=np.float32)
  ra = prog.run_host({'t1': a})['t0'][inside]
  rb = prog.run_host({'t1': b})['t0'][inside]
  rab = prog.run_host({'t1': a + b})['t0'][inside]
  assert np.max(np.abs(ra + rb - rab)) < 1e-5
  # and an exact check on a random 256 x 256 window via its dependency cone
  y0, x0, r = 1000, 2000, 8
  sub = a[y0 - r:y0 + 256 + r, x0 - r:x0 + 256 + r].copy()
  want = common.oracle_outputs(st, {'t1': sub})['t0'][r:-r, r:-r]
  got = prog.run_host({'t1': a})['t0'][y0:y0 + 256, x0:x0 + 256]
  assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize('name,extent', [('jacobi2d', (777, 300)),
                                         ('seidel2d', (500, 200)),
                                         ('xcorr', (600, 100)),
                                         ('heat3d', (200, 40, 30)),
                                         ('jacobi3d', (150, 30, 40))])
def test_computation_reuse_on_gpu(name, extent):
  """The CR-rewritten program (cr_var stages) is bit-exact against the oracle
  evaluating the same rewritten IR."""
  run_case(name, extent=extent, seed=8, computation_reuse='yes')


@pytest.mark.parametrize('name,extent,kwargs', [
    ('jacobi2d', (1000, 900), dict(iterate=11, time_block=4)),
    ('blur', (2000, 512), dict(iterate=2)),
    ('denoise2d', (500, 300), {}),
    ('heat3d', (200, 40, 90), dict(iterate=5, time_block=2)),
])
def test_pipelined_host_path(name, extent, kwargs):
  """Chunked H2D / compute / D2H overlap must not change any stored cell."""
  run_case(name, extent=extent, seed=12, host_chunks=4, **kwargs)


def test_pipelined_host_path_auto_chunks_large_grid():
  """> 32 MiB of input switches the pipeline on by default."""
  st = common.stencil('jacobi2d', iterate=8)
  prog = cuda_backend.compile_stencil(st, time_block=4)
  extent = (4096, 4096)
  rng = np.random.default_rng(4)
  a = rng.random(extent[::-1], dtype=np.float32)
  before = prog.launch_count()
  piped = prog.run_host({'t1': a})['t0']
  assert prog.launch_count() - before > prog.num_passes  # several chunks ran
  plain = prog.run_host({'t1': a}, opts=launcher.make_opts(host_chunks=1))['t0']
  assert np.array_equal(piped.view(np.uint32), plain.view(np.uint32))


def test_packed_and_scalar_paths_agree():
  """jacobi2d / heat3d use packed fp32 pairs (FADD2/FMUL2) by default; the
  scalar path (--cuda-no-pack) must give the same bits."""
  run_case('jacobi2d', extent=(1000, 700), seed=13, iterate=10, time_block=5)
  run_case('jacobi2d', extent=(1000, 700), seed=13, iterate=10, time_block=5,
           options={'no_pack': True})
  run_case('heat3d', extent=(200, 40, 50), seed=13, iterate=4, time_block=2,
           options={'no_pack': True})


def test_fast_fp_mode_tolerance():
  """--cuda-fast-fp lets ptxas contract a*b+c into FMA.  BASELINE north star:
  "<= 2 ulp per step"; it must also pass the reference's own criterion
  (abs or rel error <= 1e-5, src/soda/codegen/frt/host.py:633-649)."""
  from oracle import compare
  st = common.stencil('heat3d', iterate=4)
  prog = cuda_backend.compile_stencil(st, time_block=2,
                                      options={'fast_fp': True})
  assert prog.info.strict_fp == 0
  extent = (200, 40, 50)
  inputs = common.make_inputs(st, extent, seed=21)
  got = prog.run_host(inputs)['out']
  want = common.oracle_outputs(st, inputs)['out']
  inside = common.box_index(st.valid_box('out', extent))
  assert compare.error_count(got[inside], want[inside]) == 0
  assert compare.ulp_distance(got[inside], want[inside]).max() <= 2 * st.iterate


@pytest.mark.parametrize('name,extent', [('denoise2d', (500, 300)),
                                         ('denoise3d', (250, 37, 41))])
def test_float_math_mode(name, extent):
  """--math-precision float (sqrt(float) stays float): sqrt.rn.f32 on the GPU
  is the correctly rounded std::sqrt(float) of the oracle."""
  run_case(name, extent=extent, seed=31, math_precision='float')


def test_one_dimensional_program():
  """tests/src_extra/smooth1d.soda: a 1-D program runs lifted to N x 1
  (optimization/lift.py); bit-exact against the 1-D golden loops."""
  import os
  from soda_b200 import sodac
  with open(os.path.join(common.ROOT, 'tests', 'src_extra',
                         'smooth1d.soda')) as fp:
    st = sodac.compile_source(fp.read())
  prog = cuda_backend.compile_stencil(st)
  assert (prog.dim, prog.source_dim) == (2, 1)
  for n in (1000003, 100, 6):
    x = np.random.default_rng(n).random((n,), dtype=np.float32)
    want = common.oracle_outputs(st, {'a': x})['b']
    out = np.full((n,), 77, dtype=np.float32)
    prog.run_host({'a': x}, {'b': out})
    (lo, hi), = st.valid_box('b', (n,))
    if hi > lo:
      assert np.array_equal(out[lo:hi].view(np.uint32),
                            want[lo:hi].view(np.uint32))
    assert np.all(out[:lo] == 77) and np.all(out[max(hi, lo):] == 77)
