"""nvcc driver: generated program text -> in-tree shared library for sm_100a."""
import hashlib
import os
import shutil
import subprocess
import threading
from typing import Dict, List, Optional

from soda_b200.codegen.cuda import emit

PACKAGE_DIR = os.path.dirname(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC_DIR = os.path.join(PACKAGE_DIR, 'csrc')
INCLUDE_DIR = os.path.join(os.path.dirname(PACKAGE_DIR), 'include')
BUILD_DIR = os.path.join(PACKAGE_DIR, '_build')

ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']

# two threads asking for the same library (identical generated source) must not
# write its .cu / run nvcc on it at the same time
_LOCKS_GUARD = threading.Lock()
_LOCKS: Dict[str, threading.Lock] = {}


def _lock_for(path: str) -> threading.Lock:
  with _LOCKS_GUARD:
    return _LOCKS.setdefault(path, threading.Lock())


def nvcc_path() -> str:
  path = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
  if not os.path.exists(path):
    raise RuntimeError('nvcc not found: the CUDA backend needs the CUDA toolkit')
  return path


def nvcc_flags(strict_fp: bool = True, extra: Optional[List[str]] = None):
  flags = ['-std=c++17', '-O3', '-lineinfo', '-shared', '-Xcompiler', '-fPIC',
           # several program libraries share one process: keep template
           # statics and kernels private to each library
           '-Xcompiler', '-fvisibility=hidden', '-Xcompiler', '-fno-gnu-unique'
          ] + ARCH_FLAGS
  # the reference's results are those of g++ without FMA contraction; keep
  # float arithmetic un-contracted unless the user opts out
  flags.append('--fmad=false' if strict_fp else '--fmad=true')
  flags += ['-I', CSRC_DIR, '-I', INCLUDE_DIR]
  return flags + list(extra or [])


def _headers_digest() -> str:
  digest = hashlib.sha1()
  for directory in (CSRC_DIR, INCLUDE_DIR):
    for name in sorted(os.listdir(directory)):
      if name.endswith(('.cuh', '.h')):
        with open(os.path.join(directory, name), 'rb') as fp:
          digest.update(fp.read())
  return digest.hexdigest()


def library_path(stencil, source: str, strict_fp: bool) -> str:
  digest = hashlib.sha1(
      (source + _headers_digest() + str(strict_fp)).encode()).hexdigest()[:12]
  return os.path.join(BUILD_DIR,
                      'libsoda_%s_%s.so' % (stencil.app_name, digest))


def build_library(stencil,
                  time_block: Optional[int] = None,
                  options: Optional[Dict] = None,
                  output: Optional[str] = None,
                  keep_source: Optional[str] = None,
                  verbose: bool = False) -> str:
  """Generates and compiles the program; returns the path of the .so.

  Libraries are cached in ``soda_b200/_build`` keyed by the hash of the
  generated source, the template headers and the FP mode (the reference caches
  its slow build products by ``str(stencil)`` in the same spirit, reference:
  src/soda/optimization/cluster.py:111-121).
  """
  options = dict(options or {})
  strict_fp = not options.get('fast_fp')
  source = emit.emit_program(stencil, time_block=time_block, options=options)
  os.makedirs(BUILD_DIR, exist_ok=True)
  lib = output or library_path(stencil, source, strict_fp)
  src_path = keep_source or (os.path.splitext(lib)[0] + '.cu')
  with _lock_for(lib):
    if output is None and os.path.exists(lib):
      return lib
    # other processes (ranks of one job) may build the same library: the source
    # appears atomically, so nobody compiles a half-written file
    src_tmp = '%s.%d.%d.tmp' % (src_path, os.getpid(), threading.get_ident())
    with open(src_tmp, 'w') as fp:
      fp.write(source)
    os.replace(src_tmp, src_path)
    tmp = '%s.%d.%d.tmp' % (lib, os.getpid(), threading.get_ident())
    cmd = [nvcc_path()] + nvcc_flags(strict_fp) + ['-o', tmp, src_path]
    if verbose:
      cmd.insert(1, '-Xptxas')
      cmd.insert(2, '-v')
    result = subprocess.run(cmd, capture_output=True, text=True)
    if result.returncode != 0:
      raise RuntimeError('nvcc failed:\n%s\n%s' %
                         (' '.join(cmd), result.stderr))
    if verbose:
      print(result.stderr)
    os.replace(tmp, lib)
  return lib


def build_layout_library(verbose: bool = False) -> str:
  """Compiles the stream-data-layout codec (csrc/soda_layout.cu, C ABI
  include/soda_layout.h) for sm_100a; returns the path of the .so."""
  source = os.path.join(CSRC_DIR, 'soda_layout.cu')
  digest = hashlib.sha1()
  for path in (source, os.path.join(INCLUDE_DIR, 'soda_layout.h')):
    with open(path, 'rb') as fp:
      digest.update(fp.read())
  os.makedirs(BUILD_DIR, exist_ok=True)
  lib = os.path.join(BUILD_DIR,
                     'libsoda_layout_%s.so' % digest.hexdigest()[:12])
  if os.path.exists(lib):
    return lib
  tmp = '%s.%d.tmp' % (lib, os.getpid())
  cmd = [nvcc_path()] + nvcc_flags() + ['-o', tmp, source]
  if verbose:
    cmd[1:1] = ['-Xptxas', '-v']
  result = subprocess.run(cmd, capture_output=True, text=True)
  if result.returncode != 0:
    raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), result.stderr))
  if verbose:
    print(result.stderr)
  os.replace(tmp, lib)
  return lib
