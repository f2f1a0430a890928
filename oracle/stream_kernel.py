"""TEST INFRASTRUCTURE (oracle) - never imported by the product path.

A g++-compiled *stream kernel* with the reference kernel's C interface:

  extern "C" void <app>_kernel(ap_uint<BW>* bank_<b>_<out> ...,   // outputs first
                               ap_uint<BW>* bank_<b>_<in> ...,    // then inputs
                               uint64_t coalesced_data_num);

(reference: src/soda/codegen/xilinx/hls_kernel.py:30-202, port order
:62-66,88; called from src/soda/codegen/frt/host.py:282-289 with
``cycle_count`` of :270-276).  What the FPGA dataflow kernel computes, seen from
its ports, is a function of one-dimensional streams: element ``t`` of the
tiled input stream (docs/data-layout.md) arrives at time ``t``, a stage that
stores cell ``q`` reads its parents at ``q + serialize(ld_idx - st_idx)`` -
``serialize`` with the kernel's tile sizes, src/soda/util.py:9-12 - and the
output cell ``q`` leaves at time ``q + stencil_offset`` where ``stencil_offset =
stencil_distance - serialize(window offset)`` (the un-tiler reads it there,
src/soda/codegen/frt/host.py:396-412).  The kernel does not know tiles or image
borders: cells whose window wraps around a tile row or crosses a tile boundary
come out wrong, which is exactly why the host replicates halos and only
un-tiles the valid interior.

This file prints that function as C++ (expressions by oracle/emit_cpp.py's own
printer), so that the whole reference data path can be exercised end to end:

  dense arrays --tiler--> bank buffers --<app>_kernel--> bank buffers
              --un-tiler--> dense arrays  ==  golden loops (inside valid boxes)

with either the oracle's tiler / un-tiler (oracle/stream_layout.py) or the
CUDA pack / unpack kernels on the two ends.  It is *not* a restatement of the
reference's module / FIFO micro-architecture (forward modules, reuse chains,
unrolled PEs: those describe how an FPGA computes the same stream function and
have no observable effect at the ports); the reference cannot print its own
kernel here because ``haoda`` is absent (oracle/README.md).
"""
import ctypes
import hashlib
import os
import subprocess
from typing import Dict, List, Sequence

import numpy as np

from oracle import emit_cpp
from soda_b200 import core, util

BUILD_DIR = emit_cpp.BUILD_DIR


def emit(stencil) -> str:
  """C++ source of ``<app>_kernel`` for ``stencil`` (its own tile sizes)."""
  dim = stencil.dim
  tile = tuple(stencil.tile_size)
  tensors = stencil.chronological_tensors
  burst = stencil.burst_width
  lines = [emit_cpp._PRELUDE]
  lines.append('// %s' % str(stencil).replace('\n', '\n// '))
  ports = []
  for stmt in stencil.output_stmts + stencil.input_stmts:
    for bank in stmt.dram:
      ports.append('void* bank_%d_%s' % (bank, stmt.name))
  lines.append('extern "C" void %s_kernel(%s, uint64_t coalesced_data_num) {' %
               (stencil.app_name, ', '.join(ports)))

  # stream lengths: every bank carries coalesced_data_num words of
  # burst_width bits (host.py:270-276); elements are dealt to the banks of a
  # tensor cyclically (docs/data-layout.md, host.py:243-245)
  def per_word(stmt):
    return burst // stmt.haoda_type.width_in_bits

  first = stencil.input_stmts[0]
  lines.append('  const int64_t total = int64_t(coalesced_data_num) * %d * %d;' %
               (per_word(first), len(first.dram)))
  for stmt in stencil.input_stmts:
    ctype = emit_cpp._ctype(stmt.haoda_type)
    banks = len(stmt.dram)
    lines.append('  std::vector<%s> data_%s(total);' % (ctype, stmt.name))
    lines.append('  {')
    lines.append('    const %s* banks[%d] = {%s};' % (ctype, banks, ', '.join(
        'static_cast<const %s*>(bank_%d_%s)' % (ctype, b, stmt.name)
        for b in stmt.dram)))
    lines.append('    const int64_t have = int64_t(coalesced_data_num) * %d * %d;'
                 % (per_word(stmt), banks))
    lines.append('    for (int64_t t = 0; t < total; ++t)')
    lines.append('      data_%s[t] = t < have ? banks[t %% %d][t / %d] : %s(0);' %
                 (stmt.name, banks, banks, ctype))
    lines.append('  }')

  for tensor in tensors:
    if tensor.is_input():
      continue
    ctype = emit_cpp._ctype(tensor.haoda_type)
    lines.append('')
    lines.append('  // stage %s: cell q reads its parents at q + serialize(ld - st)'
                 % tensor.name)
    lines.append('  std::vector<%s> data_%s(total);' % (ctype, tensor.name))
    lines.append('  for (int64_t q = 0; q < total; ++q) {')
    st_idx = tensor.st_idx

    def load(ref, st_idx=st_idx):
      if ref.name in stencil.param_names:
        raise util.SemanticError('stream kernels with params are not printed')
      delta = tuple(a - b for a, b in zip(ref.idx, st_idx))
      offset = util.serialize(delta, tile)
      return 'at(data_%s, q + (%d), total)' % (ref.name, offset)

    variables = {}
    printer = emit_cpp._Expr(load, variables)
    for let in tensor.lets:
      text, source = printer(let.expr)
      t = let.haoda_type if let.haoda_type is not None else source
      lines.append('    const %s %s = %s;' % (emit_cpp._ctype(t), let.name,
                                               emit_cpp._convert(t, text, source)))
      variables[let.name] = t
    text, source = printer(tensor.expr)
    lines.append('    data_%s[q] = %s;' %
                 (tensor.name, emit_cpp._convert(tensor.haoda_type, text, source)))
    lines.append('  }')

  for stmt in stencil.output_stmts:
    tensor = stencil.tensors[stmt.name]
    ctype = emit_cpp._ctype(stmt.haoda_type)
    banks = len(stmt.dram)
    # host.py:396-403
    window = core.get_overall_stencil_window(
        [stencil.tensors[n] for n in stencil.input_names], tensor)
    distance = core.get_stencil_distance(window, stencil.tile_size)
    offset = distance - util.serialize(core.get_stencil_window_offset(window),
                                       stencil.tile_size)
    lines.append('')
    lines.append('  // output %s: cell q leaves at time q + %d' % (stmt.name,
                                                                   offset))
    lines.append('  {')
    lines.append('    %s* banks[%d] = {%s};' % (ctype, banks, ', '.join(
        'static_cast<%s*>(bank_%d_%s)' % (ctype, b, stmt.name)
        for b in stmt.dram)))
    lines.append('    const int64_t have = int64_t(coalesced_data_num) * %d * %d;'
                 % (per_word(stmt), banks))
    lines.append('    for (int64_t s = 0; s < have; ++s)')
    lines.append('      banks[s %% %d][s / %d] = at(data_%s, s - (%d), total);' %
                 (banks, banks, stmt.name, offset))
    lines.append('  }')
  lines.append('}')
  helper = ('namespace {\ntemplate <typename T> inline T at(const std::vector<T>& '
            'v, int64_t i, int64_t n) {\n  return i >= 0 && i < n ? v[i] : T(0);'
            '  // before the stream starts / after it ends\n}\n}  // namespace\n')
  source = '\n'.join(lines) + '\n'
  marker = '// kernel:'
  head, _, tail = source.partition(marker)
  return head + helper + marker + tail


def build(stencil) -> str:
  source = emit(stencil)
  digest = hashlib.sha1(source.encode()).hexdigest()[:12]
  os.makedirs(BUILD_DIR, exist_ok=True)
  base = os.path.join(BUILD_DIR, 'stream_%s_%s' % (stencil.app_name, digest))
  lib = base + '.so'
  if os.path.exists(lib):
    return lib
  with open(base + '.cpp', 'w') as fp:
    fp.write(source)
  tmp = '%s.%d.tmp.so' % (base, os.getpid())
  subprocess.run(['g++', '-std=c++17', '-shared', '-fPIC'] +
                 emit_cpp.PARITY_FLAGS + [base + '.cpp', '-o', tmp], check=True)
  os.replace(tmp, lib)
  return lib


def cycle_count(stencil, extent: Sequence[int]) -> int:
  """``cycle_count`` of the reference host (src/soda/codegen/frt/host.py:270-276)
  for the program's own tile sizes and burst width."""
  dim = stencil.dim
  first = stencil.input_stmts[0]
  epc = stencil.burst_width // first.haoda_type.width_in_bits * len(first.dram)
  stencil_dim = core.get_stencil_dim(stencil.stencil_window)
  count = extent[dim - 1]
  for d in range(dim - 1):
    count *= stencil.tile_size[d]
    count *= (extent[d] - stencil_dim[d] + 1 - 1) // (
        stencil.tile_size[d] - stencil_dim[d] + 1) + 1
  return (count + stencil.stencil_distance - 1) // epc + 1


class StreamKernel:
  """ctypes wrapper: bank buffers in, bank buffers out."""

  def __init__(self, stencil):
    self.stencil = stencil
    self.lib = ctypes.CDLL(build(stencil))
    self.func = getattr(self.lib, stencil.app_name + '_kernel')
    self.func.restype = None

  def run(self, in_banks: Dict[str, List[np.ndarray]],
          out_banks: Dict[str, List[np.ndarray]], cycles: int) -> None:
    args = []
    keep = []
    for stmt in self.stencil.output_stmts:
      for array in out_banks[stmt.name]:
        assert array.flags.c_contiguous and array.flags.writeable
        args.append(ctypes.c_void_p(array.ctypes.data))
    for stmt in self.stencil.input_stmts:
      for array in in_banks[stmt.name]:
        array = np.ascontiguousarray(array)
        keep.append(array)
        args.append(ctypes.c_void_p(array.ctypes.data))
    args.append(ctypes.c_uint64(cycles))
    self.func(*args)
