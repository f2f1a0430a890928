"""Differential testing on seeded random programs (tests/random_programs.py).

Three evaluators per program: NumPy golden loops, the g++ oracle, and the CUDA
templates - emulated on the CPU here, on the GPU with ``-m gpu`` (libraries
prebuilt by __graft_entry__.build()).  Bit-exact inside the valid box, border
untouched.  The hand-written programs of tests/src pin known shapes; these
pin the planner (lags, halos, window depths, store boxes, shuffles across
lane boundaries, shared-memory reach) on shapes nobody chose."""
import numpy as np
import pytest

from oracle import emit_cpp, golden
from soda_b200 import sodac
from soda_b200.codegen.cuda import launcher
from tests import common, random_programs
from tests.emu import build_emu

CPU_SEEDS = list(range(12))
GPU_SEEDS = list(range(40))


def _case(seed):
  text, extent, kwargs = random_programs.program(seed)
  st = sodac.compile_source(text)
  return st, extent, kwargs, random_programs.inputs_for(st, extent, seed)


def _check(st, extent, prog, inputs):
  want = emit_cpp.Oracle(st).run(inputs)
  dtype = golden.np_dtype(st.output_stmts[0].haoda_type)
  outputs = {'out': np.full(extent[::-1], 77, dtype=dtype)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)


def test_generator_is_deterministic_and_covers_the_families():
  assert random_programs.program(5) == random_programs.program(5)
  seen = set()
  for seed in GPU_SEEDS:
    st, _, kwargs, _ = _case(seed)
    seen.add((st.dim, str(st.input_types[0])))
    assert str(sodac.compile_source(str(st))) == str(st)  # text round trip
  assert {t for _, t in seen} >= {'float', 'half', 'int16', 'int32', 'uint16',
                                  'uint8', 'uint6'}
  assert {d for d, _ in seen} == {2, 3}


@pytest.mark.parametrize('seed', GPU_SEEDS)
def test_oracles_agree(seed):
  st, extent, _, inputs = _case(seed)
  a = golden.run(st, inputs)['out']
  b = emit_cpp.Oracle(st).run(inputs)['out']
  index = common.box_index(st.valid_box('out', extent))
  assert a[index].size > 0
  assert np.array_equal(a[index].view(np.uint8), b[index].view(np.uint8))
  assert np.isfinite(a[index].astype(np.float64)).all()


@pytest.mark.parametrize('seed', CPU_SEEDS)
def test_under_emulation(seed):
  st, extent, kwargs, inputs = _case(seed)
  prog = launcher.CudaProgram(build_emu.build_emu_library(st, **kwargs))
  _check(st, extent, prog, inputs)


@pytest.mark.gpu
@pytest.mark.parametrize('seed', GPU_SEEDS)
def test_on_gpu(seed):
  from soda_b200.codegen import cuda as cuda_backend
  st, extent, kwargs, inputs = _case(seed)
  _check(st, extent, cuda_backend.compile_stencil(st, **kwargs), inputs)


# ---- computation reuse on random sum stencils --------------------------------------

@pytest.mark.parametrize('seed', range(16))
def test_computation_reuse_keeps_the_value(seed):
  """Every scheduler (native soda-cr contract, greedy, built-in) must return a
  program that computes the same sums with no more additions: exactly for
  integer tensors, within rounding for float ones."""
  text, extent, taps = random_programs.sum_program(seed)
  plain = sodac.compile_source(text)
  inputs = random_programs.inputs_for(plain, extent, seed)
  want = golden.run(plain, inputs)['b']
  index = common.box_index(plain.valid_box('b', extent))
  for mode in ('yes', 'greedy', 'built-in'):
    reused = sodac.compile_source(text, computation_reuse=mode)
    assert reused.valid_box('b', extent) == plain.valid_box('b', extent)
    assert str(reused).count(' + ') <= taps - 1
    got = golden.run(reused, inputs)['b']
    if plain.input_types[0].is_float:
      assert np.allclose(want[index], got[index], rtol=1e-5, atol=1e-6)
    else:
      assert np.array_equal(want[index], got[index])


@pytest.mark.parametrize('seed', [0, 3, 8, 16])
def test_computation_reuse_under_emulation(seed):
  """The rewritten program (shared partial sums as fused local stages) through
  the CUDA templates, bit-exact against its own golden loops."""
  text, extent, _ = random_programs.sum_program(seed)
  st = sodac.compile_source(text, computation_reuse='yes')
  inputs = random_programs.inputs_for(st, extent, seed)
  prog = launcher.CudaProgram(build_emu.build_emu_library(st))
  dtype = golden.np_dtype(st.output_stmts[0].haoda_type)
  outputs = {'b': np.full(extent[::-1], 77, dtype=dtype)}
  prog.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs,
                               emit_cpp.Oracle(st).run(inputs), sentinel=77)
