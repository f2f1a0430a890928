"""GPU parity at BASELINE.json's full sizes, through the C ABI, for exactly the
libraries bench.py measures (same config table, planner defaults).

A full golden run of 16384^2 x 64 or 512^3 x 32 takes minutes on the CPU, so
the measured output is checked on seeded windows - the corners of the valid
box, random positions - bit for bit against the g++ oracle run on each
window's dependency cone (oracle/cone.py; reference semantics:
src/soda/codegen/frt/host.py:556-624)."""
import numpy as np
import pytest
import torch

import bench
from oracle import cone
from soda_b200.codegen import cuda as cuda_backend
from soda_b200.codegen.cuda import launcher

pytestmark = pytest.mark.gpu


def host_inputs(st, prog, extent):
  cpu = torch.device('cpu')
  box = tuple((0, e) for e in extent)
  return {
      name: bench.synthetic(box, index, getattr(torch, str(dtype)),
                            cpu).numpy()
      for index, (name, dtype) in enumerate(zip(prog.input_names,
                                                prog.input_dtypes))
  }


def check(st, prog, extent, windows=18, opts=None):
  inputs = host_inputs(st, prog, extent)
  outputs = {n: np.full(extent[::-1], 77, dtype=d)
             for n, d in zip(prog.output_names, prog.output_dtypes)}
  before = prog.launch_count()
  prog.run_host(inputs, outputs, opts=opts)
  assert prog.launch_count() > before, 'no kernel was launched'
  report = cone.check_host_arrays(st, inputs, outputs, count=windows, seed=2)
  assert report['windows'] >= windows
  assert report['bit_exact'], report['first_mismatch']
  # cells outside the valid box keep the caller's bytes
  for name in prog.output_names:
    box = st.valid_box(name, extent)
    array = outputs[name]
    for d, (lo, hi) in enumerate(box):
      axis = array.ndim - 1 - d
      if lo > 0:
        assert np.all(np.take(array, range(0, lo), axis=axis) == 77)
      if hi < extent[d]:
        assert np.all(np.take(array, range(hi, extent[d]), axis=axis) == 77)
  return report


def test_bench_headline_library_c2():
  """jacobi2d 16384^2 iterate 64 with the planner's time block and launch
  shape, measured segment length on: the very library bench.py times."""
  st, prog = bench.config_program('C2_jacobi2d')
  assert prog.num_passes == -(-64 // prog.pass_info(0).time_block)
  check(st, prog, bench.CONFIGS['C2_jacobi2d']['extent'])


def test_bench_headline_library_c2_host_pipeline_seams():
  """The chunked H2D / compute / D2H pipeline of the e2e number: windows on
  the seams between 16 equal chunks."""
  st, prog = bench.config_program('C2_jacobi2d')
  extent = bench.CONFIGS['C2_jacobi2d']['extent']
  inputs = host_inputs(st, prog, extent)
  outputs = {'t0': np.zeros(extent[::-1], dtype=np.float32)}
  prog.run_host(inputs, outputs, opts=launcher.make_opts(host_chunks=16))
  seams = [(int(x), extent[1] * k // 16 - 32)
           for k, x in zip(range(1, 16), np.linspace(100, 16000, 15))]
  report = cone.check_host_arrays(st, inputs, outputs, count=2, seed=4,
                                  required=seams)
  assert report['windows'] == 17 and report['bit_exact'], report


def automatic_chunk_bounds(slices, reach, chunks=20, ratio=0.92):
  """The bounds HostPipeline::issue() (soda_runtime.cuh) gives the automatic
  layout of a grid this large: lengths in geometric progression, in steps of
  64 slices, none shorter than the reach of all passes."""
  total = sum(ratio ** k for k in range(chunks))
  step = 64 if slices >= 64 * 4 * chunks else 1
  shortest = max(step, reach)
  bounds, run = [0], 0.0
  for k in range(chunks):
    run += ratio ** k
    b = int(slices * (run / total) / step + 0.5) * step
    b = max(b, bounds[k] + shortest)
    b = min(b, slices - (chunks - 1 - k) * shortest)
    bounds.append(b)
  bounds[-1] = slices
  return bounds


def test_bench_headline_library_c2_host_pipeline_automatic_layout():
  """The layout the e2e number runs with (chunk count 0: 20 chunks that
  shrink towards the end): windows on every seam."""
  st, prog = bench.config_program('C2_jacobi2d')
  extent = bench.CONFIGS['C2_jacobi2d']['extent']
  inputs = host_inputs(st, prog, extent)
  outputs = {'t0': np.zeros(extent[::-1], dtype=np.float32)}
  prog.run_host(inputs, outputs, opts=launcher.make_opts(host_chunks=0))
  bounds = automatic_chunk_bounds(extent[1], 2 * 64)
  assert len(bounds) == 21 and bounds[1] - bounds[0] > 4 * (bounds[-1] - bounds[-2])
  seams = [(int(x), b - 32)
           for b, x in zip(bounds[1:-1], np.linspace(100, 16000, len(bounds) - 2))]
  report = cone.check_host_arrays(st, inputs, outputs, count=2, seed=5,
                                  required=seams)
  assert report['windows'] == len(seams) + 2 and report['bit_exact'], report


def test_c1_blur_2000_wide():
  st, prog = bench.config_program('C1_blur')
  check(st, prog, bench.CONFIGS['C1_blur']['extent'])


@pytest.mark.parametrize('key', ['C3_heat3d', 'C3_jacobi3d'])
@pytest.mark.parametrize('time_block', [None, 1, 2])
def test_c3_512_cubed(key, time_block):
  st = bench.config_stencil(key)
  prog = cuda_backend.compile_stencil(st, time_block=time_block)
  check(st, prog, bench.CONFIGS[key]['extent'], windows=12)


def test_c3_heat3d_with_fused_power_of_two_coefficients():
  """--cuda-pow2-fma: `c * x + acc` with a power-of-two literal c in one FFMA;
  the product is exact, so on data in the normal range the result is the
  un-fused one, bit for bit."""
  st, prog = bench.config_program('C3_heat3d_pow2_fma')
  check(st, prog, bench.CONFIGS['C3_heat3d_pow2_fma']['extent'], windows=12)


@pytest.mark.parametrize('key', ['C4_denoise3d', 'C4_denoise3d_cr'])
def test_c4_denoise3d_512_cubed(key):
  st, prog = bench.config_program(key)
  check(st, prog, bench.CONFIGS[key]['extent'], windows=12)
