"""The C ABI from a plain C caller (gcc, no Python in the data path): the
binding INTEGRATION.md section 2 describes.  The CPU test compiles and links
tests/c_host/jacobi2d_host.c against a program library; the `gpu` test runs it
on B200 and compares with the oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from soda_b200.codegen.cuda import build
from tests import common

HOST_C = os.path.join(common.ROOT, 'tests', 'c_host', 'jacobi2d_host.c')
INCLUDE = os.path.join(common.ROOT, 'include')


def build_host(tmp_path):
  st = common.stencil('jacobi2d')
  lib = build.build_library(st)
  exe = os.path.join(str(tmp_path), 'jacobi2d_host')
  subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-O1', '-I', INCLUDE,
                  HOST_C, lib, '-Wl,-rpath,' + os.path.dirname(lib), '-o', exe],
                 check=True)
  return st, exe


def test_c_host_compiles_and_links(tmp_path):
  _, exe = build_host(tmp_path)
  assert os.access(exe, os.X_OK)
  # without arguments it prints its usage and exits before touching CUDA
  result = subprocess.run([exe], capture_output=True, text=True)
  assert result.returncode == 2 and 'usage' in result.stderr


@pytest.mark.gpu
def test_c_host_matches_the_oracle(tmp_path):
  st, exe = build_host(tmp_path)
  extent = (333, 77)
  inputs = common.make_inputs(st, extent, seed=11)
  in_path = os.path.join(str(tmp_path), 'in.bin')
  out_path = os.path.join(str(tmp_path), 'out.bin')
  inputs['t1'].tofile(in_path)
  np.full(extent[::-1], 77, dtype=np.float32).tofile(out_path)
  result = subprocess.run([exe, str(extent[0]), str(extent[1]), in_path,
                           out_path], capture_output=True, text=True)
  assert result.returncode == 0, result.stderr
  assert 'program jacobi2d: 2-D, iterate 2' in result.stdout
  got = np.fromfile(out_path, dtype=np.float32).reshape(extent[::-1])
  want = common.oracle_outputs(st, inputs)
  common.assert_matches_oracle(st, extent, {'t0': got}, want, sentinel=77)
