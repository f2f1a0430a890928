"""TEST INFRASTRUCTURE (oracle) - never imported by the product path.

The reference-style *dataflow* kernel, printed as HLS C++ and compiled with g++
against the stand-in headers in oracle/shim/ (`hls_stream.h`, `ap_int.h`): the
second, structure-faithful CPU baseline of SURVEY section 8(f) row 4.

What the reference prints (src/soda/codegen/xilinx/hls_kernel.py:338-443 with
the graph of src/soda/dataflow.py:336-625) and this module restates from this
repo's own IR:

* the module graph: per input bank a *load* (unpack) module, per producer
  tensor and accessed stream offset a *forward* module (a delay line of
  ``reuse_buffer_lengths`` elements, chained along the reuse chain of its
  unroll lane), per stage and unroll lane a *compute* module (one processing
  element), per output bank a *store* (pack) module
  (dataflow.py:346-355, :448-514; offsets and chains are
  ``Stencil.all_points / reuse_buffers / next_fifo``, core.py:505-563,684-795);
* tokens ``Data<T>{data, ctrl}`` through ``hls::stream`` FIFOs, one FIFO per
  edge, a module loop ``for (bool enable = true; enable;)`` that ends with the
  token whose ``ctrl`` is false (hls_kernel.py:721-774, :901-939);
* delay lines as circular buffers with a pointer (``ir.DelayedRef``,
  hls_kernel.py:711-718,785-786,879-880);
* ``BurstRead`` / ``BurstWrite`` between the ``ap_uint<burst width>*`` ports
  and burst-word FIFOs, slicing of a burst word into elements with the range
  operator, cyclic over the banks of a tensor (hls_kernel.py:238-263,
  :444-491, :788-869);
* the kernel top ``extern "C" void <app>_kernel(outputs..., inputs...,
  uint64_t coalesced_data_num)`` wiring everything inside one dataflow region
  (hls_kernel.py:30-202); identical module bodies are printed once as
  ``Module<N>Func`` (dataflow.py:185-202).

The reference needs ``haoda`` (module / FIFO IR and the C printer) and the
Xilinx headers to do this; neither exists here, so the text below is this
repo's own and only its *architecture* is the reference's.  Under C simulation
the modules of a dataflow region run one after the other over unbounded
streams; oracle/shim/hls_stream.h is exactly that.

Verified by tests/test_dataflow_kernel.py: host tiler -> this kernel -> host
un-tiler equals the golden loops inside every valid box (the harness of
oracle/stream_kernel.py, which prints the same function of the streams without
any micro-architecture; the two differ only in void cells, where the delay
lines start from zeros instead of a zero-extended stream).

Only tests/ and bench.py's `--impl reference-dataflow` leg may execute it.
"""
import collections
import ctypes
import hashlib
import os
import subprocess
from typing import Dict, List, Tuple

import numpy as np

from oracle import emit_cpp
from oracle.stream_kernel import cycle_count  # noqa: F401  (same contract)
from soda_b200 import util

BUILD_DIR = emit_cpp.BUILD_DIR
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


class Fifo:
  """One edge of the module graph."""

  def __init__(self, src: 'Module', dst: 'Module', ctype: str):
    self.src, self.dst, self.ctype = src, dst, ctype
    self.name = 'from_%s_to_%s' % (src.name, dst.name)


class Module:
  """kind: load | forward | compute | store."""

  def __init__(self, kind: str, name: str):
    self.kind = kind
    self.name = name
    self.inputs: List[Fifo] = []
    self.outputs: List[Fifo] = []
    # load / store
    self.var = None
    self.bank = None        # DRAM bank number (port name)
    self.positions = ()     # unroll positions served by this bank
    # forward
    self.delay = 0
    # compute
    self.tensor = None
    self.pe = None
    self.ref_fifo: Dict[Tuple[str, Tuple[int, ...]], int] = {}


def _connect(src: Module, dst: Module, ctype: str) -> Fifo:
  fifo = Fifo(src, dst, ctype)
  src.outputs.append(fifo)
  dst.inputs.append(fifo)
  return fifo


def build_graph(stencil) -> List[Module]:
  """Modules in an order in which every module comes after its producers."""
  uf = stencil.unroll_factor
  if stencil.param_stmts:
    raise util.SemanticError('dataflow kernels with params are not printed')
  tensors = stencil.tensors
  all_points = stencil.all_points
  reuse_buffers = stencil.reuse_buffers
  lengths = stencil.reuse_buffer_lengths
  modules: List[Module] = []
  stmt_of = {s.name: s for s in stencil.input_stmts + stencil.output_stmts}

  def check_banks(stmt):
    banks = len(stmt.dram)
    if uf % banks:
      raise util.SemanticError(
          'unroll factor %d is not a multiple of the %d banks of %s' %
          (uf, banks, stmt.name))
    return banks

  # load modules: bank b of an input carries the unroll positions k with
  # k % banks == index of b (elements are dealt to the banks cyclically)
  loads: Dict[Tuple[str, int], Module] = {}
  for stmt in stencil.input_stmts:
    banks = check_banks(stmt)
    for index, bank in enumerate(stmt.dram):
      module = Module('load', '%s_bank_%d' % (stmt.name, bank))
      module.var, module.bank = stmt.name, bank
      module.positions = tuple(k for k in range(uf) if k % banks == index)
      loads[stmt.name, index] = module
      modules.append(module)

  computes: Dict[Tuple[str, int], Module] = {}
  forwards: Dict[Tuple[str, int], Module] = {}

  def add_forward_modules(tensor):
    """The reuse chains of a producer tensor, one per unroll lane."""
    ctype = emit_cpp._ctype(tensor.haoda_type)
    chains: Dict[int, List[int]] = collections.defaultdict(list)
    for start, end in reuse_buffers[tensor.name][1:]:
      chains[end % uf].append(end)
    for lane in sorted(chains, reverse=True):
      position = uf - 1 - lane  # which element of a cycle this lane carries
      if tensor.is_input():
        banks = len(stmt_of[tensor.name].dram)
        previous = loads[tensor.name, position % banks]
      else:
        previous = computes[tensor.name, position]
      for offset in sorted(chains[lane]):
        module = Module('forward', '%s_offset_%d' % (tensor.name, offset))
        module.delay = lengths[tensor.name][offset]
        module.tensor = tensor
        _connect(previous, module, ctype)
        forwards[tensor.name, offset] = module
        modules.append(module)
        previous = module

  for tensor in stencil.chronological_tensors:
    if tensor.is_input():
      add_forward_modules(tensor)
      continue
    for pe in range(uf):
      module = Module('compute', '%s_pe_%d' % (tensor.name, pe))
      module.tensor, module.pe = tensor, pe
      for parent_name, by_offset in tensor.ld_offsets.items():
        points = all_points[parent_name][tensor.name]
        # one FIFO per distinct accessed element (a reference that appears
        # twice in the expression reads the same token)
        for index, ref in enumerate(by_offset.values()):
          offset = next(o for o, by_pe in points.items()
                        if by_pe.get(pe) == index)
          fifo = _connect(forwards[parent_name, offset], module,
                          emit_cpp._ctype(tensors[parent_name].haoda_type))
          module.ref_fifo[parent_name, tuple(ref.idx)] = \
              module.inputs.index(fifo)
      computes[tensor.name, pe] = module
      modules.append(module)
    if tensor.is_output():
      if tensor.name not in stmt_of:
        raise util.SemanticError('%s has no consumer and is not an output' %
                                 tensor.name)
    else:
      add_forward_modules(tensor)

  # store modules, after everything that feeds them
  for stmt in stencil.output_stmts:
    banks = check_banks(stmt)
    ctype = emit_cpp._ctype(stmt.haoda_type)
    for index, bank in enumerate(stmt.dram):
      module = Module('store', '%s_bank_%d' % (stmt.name, bank))
      module.var, module.bank = stmt.name, bank
      module.positions = tuple(k for k in range(uf) if k % banks == index)
      for pe in module.positions:
        _connect(computes[stmt.name, pe], module, ctype)
      modules.append(module)
  return modules


# ---- text ---------------------------------------------------------------------

_HEADER = r'''// generated by oracle/dataflow_kernel.py - test infrastructure, not product
// architecture: reference src/soda/codegen/xilinx/hls_kernel.py (m_axi interface)
#include <cfloat>
#include <cmath>
#include <cstdbool>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <algorithm>
#include <tuple>

#include <ap_int.h>
#include <hls_stream.h>

template <typename To, typename From>
To Reinterpret(From val) {
#pragma HLS inline
  static_assert(sizeof(To) == sizeof(From), "same storage");
  To out;
  std::memcpy(&out, &val, sizeof(To));
  return out;
}

template <typename T>
struct Data {
  T data;
  bool ctrl;
};

template <typename T>
bool ReadData(T& data, hls::stream<Data<T>>& from) {
#pragma HLS inline
  const auto tmp = from.read();
  data = tmp.data;
  return tmp.ctrl;
}

template <typename T>
void WriteData(hls::stream<Data<T>>& to, const T& data, bool ctrl) {
#pragma HLS inline
  Data<T> tmp;
  tmp.data = data;
  tmp.ctrl = ctrl;
  to.write(tmp);
}

template <typename T>
void BurstRead(hls::stream<Data<T>>& to, T* from, uint64_t data_num) {
load:
  for (uint64_t i = 0; i < data_num;) {
#pragma HLS pipeline II = 1
    const uint64_t next_i = i + 1;
    WriteData(to, from[i], next_i < data_num);
    i = next_i;
  }
}

template <typename T>
void BurstWrite(T* to, hls::stream<Data<T>>& from, uint64_t data_num) {
store:
  for (uint64_t i = 0; i < data_num; ++i) {
#pragma HLS pipeline II = 1
    T buf;
    ReadData(buf, from);
    to[i] = buf;
  }
}
'''

# the oracle's expression helpers (o_min, o_div, o_wrap ...) without its includes
_HELPERS = emit_cpp._PRELUDE[emit_cpp._PRELUDE.index('namespace {'):]


def _coalescing(stencil, stmt, module) -> Tuple[int, int, int]:
  """(element bits, elements of this bank per cycle, cycles per burst word)."""
  bits = stmt.haoda_type.width_in_bits
  if bits % 8:
    raise util.SemanticError('%s: %d-bit elements are not byte-sized' %
                             (stmt.name, bits))
  batch = len(module.positions)
  per_word = stencil.burst_width // bits
  if per_word % batch:
    raise util.SemanticError(
        'cannot process such a burst: %d elements per word, %d per cycle' %
        (per_word, batch))
  return bits, batch, per_word // batch


def _module_body(stencil, module: Module) -> Tuple[List[str], List[str]]:
  """(parameter declarations, body lines) with positional FIFO names
  (``fifo_st_<i>``, ``fifo_ld_<i>``), so that equal modules print equal text."""
  burst = stencil.burst_width
  stmt_of = {s.name: s for s in stencil.input_stmts + stencil.output_stmts}
  params = ['/*output*/ hls::stream<Data<%s>>& fifo_st_%d' % (f.ctype, i)
            for i, f in enumerate(module.outputs)]
  if module.kind == 'store':
    params.append('/*output*/ hls::stream<Data<ap_uint<%d>>>& dram_fifo' % burst)
  params += ['/* input*/ hls::stream<Data<%s>>& fifo_ld_%d' % (f.ctype, i)
             for i, f in enumerate(module.inputs)]
  if module.kind == 'load':
    params.append('/* input*/ hls::stream<Data<ap_uint<%d>>>& dram_fifo' % burst)

  body: List[str] = []
  if module.kind == 'forward' and module.delay > 0:
    ctype = module.inputs[0].ctype
    body += ['  uint32_t ptr_delay_%d = 0;' % module.delay,
             '  static thread_local %s buf_delay_%d[%d];' %
             (ctype, module.delay, module.delay),
             '  std::fill(buf_delay_{0}, buf_delay_{0} + {0}, {1}(0));'.format(
                 module.delay, ctype)]
  body.append('@LABEL@')
  body.append('  for (bool enable = true; enable;) {')
  body.append('#pragma HLS pipeline II = 1')
  watched = ['fifo_ld_%d' % i for i in range(len(module.inputs))]
  if module.kind == 'load':
    watched.append('dram_fifo')
  body.append('    if (%s) {' % ' && '.join('!%s.empty()' % w for w in watched))

  def read_all(prefix='      ', declare=True):
    for i, fifo in enumerate(module.inputs):
      if declare:
        body.append('%s%s fifo_ref_%d;' % (prefix, fifo.ctype, i))
      body.append('%sconst bool fifo_ref_%d_enable = ReadData(fifo_ref_%d, '
                  'fifo_ld_%d);' % (prefix, i, i, i))

  if module.kind == 'load':
    stmt = stmt_of[module.var]
    bits, batch, cycles = _coalescing(stencil, stmt, module)
    ctype = emit_cpp._ctype(stmt.haoda_type)
    body += ['      ap_uint<%d> dram_buf;' % burst,
             '      const bool dram_buf_enable = ReadData(dram_buf, dram_fifo);',
             '      const bool enabled = dram_buf_enable;',
             '      enable = enabled;']
    for cycle in range(cycles):
      for j in range(batch):
        lsb = (cycle * batch + j) * bits
        body.append(
            '      WriteData(fifo_st_%d, Reinterpret<%s>(static_cast<ap_uint<%d>>('
            'dram_buf(%d, %d))), %s);' %
            (j, ctype, bits, lsb + bits - 1, lsb,
             'true' if cycle < cycles - 1 else 'enabled'))
  elif module.kind == 'store':
    stmt = stmt_of[module.var]
    bits, batch, cycles = _coalescing(stencil, stmt, module)
    body.append('      ap_uint<%d> dram_buf;' % burst)
    for i, fifo in enumerate(module.inputs):
      body.append('      %s fifo_ref_%d;' % (fifo.ctype, i))
    for cycle in range(cycles):
      last = cycle == cycles - 1
      for i in range(batch):
        body.append('      %sReadData(fifo_ref_%d, fifo_ld_%d);' %
                    ('const bool fifo_ref_%d_enable = ' % i if last else '',
                     i, i))
      if last:
        body += ['      const bool enabled = %s;' % ' && '.join(
            'fifo_ref_%d_enable' % i for i in range(batch)),
                 '      enable = enabled;']
      for i in range(batch):
        lsb = (cycle * batch + i) * bits
        body.append('      dram_buf(%d, %d) = Reinterpret<ap_uint<%d>>('
                    'fifo_ref_%d);' % (lsb + bits - 1, lsb, bits, i))
    body.append('      WriteData(dram_fifo, dram_buf, enabled);')
  elif module.kind == 'forward':
    read_all()
    body += ['      const bool enabled = fifo_ref_0_enable;',
             '      enable = enabled;']
    ctype = module.inputs[0].ctype
    value = 'fifo_ref_0'
    if module.delay > 0:
      body.append('      const %s let_0 = buf_delay_%d[ptr_delay_%d];' %
                  (ctype, module.delay, module.delay))
      value = 'let_0'
    for i in range(len(module.outputs)):
      body.append('      WriteData(fifo_st_%d, %s(%s), enabled);' %
                  (i, ctype, value))
    if module.delay > 0:
      body += ['      buf_delay_{0}[ptr_delay_{0}] = fifo_ref_0;'.format(
          module.delay),
               '      ptr_delay_{0} = ptr_delay_{0} < {1} ? ptr_delay_{0} + 1 : 0;'
               .format(module.delay, module.delay - 1)]
  else:  # compute: one processing element
    tensor = module.tensor
    read_all()
    body += ['      const bool enabled = %s;' % ' && '.join(
        'fifo_ref_%d_enable' % i for i in range(len(module.inputs))),
             '      enable = enabled;']

    def load(ref):
      return 'fifo_ref_%d' % module.ref_fifo[ref.name, tuple(ref.idx)]

    variables = {}
    printer = emit_cpp._Expr(load, variables)
    for let in tensor.lets:
      text, source = printer(let.expr)
      t = let.haoda_type if let.haoda_type is not None else source
      body.append('      const %s %s = %s;' %
                  (emit_cpp._ctype(t), let.name,
                   emit_cpp._convert(t, text, source)))
      variables[let.name] = t
    text, source = printer(tensor.expr)
    ctype = emit_cpp._ctype(tensor.haoda_type)
    body.append('      const %s result = %s;' %
                (ctype, emit_cpp._convert(tensor.haoda_type, text, source)))
    for i in range(len(module.outputs)):
      body.append('      WriteData(fifo_st_%d, %s(result), enabled);' %
                  (i, ctype))
  body += ['    }', '  }']
  return params, body


def emit(stencil) -> str:
  """HLS C++ source of the dataflow kernel of ``stencil``."""
  modules = build_graph(stencil)
  burst = stencil.burst_width
  lines = [_HEADER, _HELPERS]
  lines.append('// %s' % str(stencil).replace('\n', '\n// '))
  lines.append('')

  # module definitions, one per distinct body
  func_of: Dict[str, int] = {}
  module_func: Dict[Module, int] = {}
  for module in modules:
    params, body = _module_body(stencil, module)
    key = '\n'.join(params + body)
    if key not in func_of:
      index = func_of[key] = len(func_of)
      lines.append('// %s module' % module.kind)
      lines.append('void Module%dFunc(\n  %s)' % (index, ',\n  '.join(params)))
      lines.append('{')
      lines += ['module_%d:' % index if l == '@LABEL@' else l for l in body]
      lines.append('}')
      lines.append('')
    module_func[module] = func_of[key]

  outputs = [(s.name, bank) for s in stencil.output_stmts for bank in s.dram]
  inputs = [(s.name, bank) for s in stencil.input_stmts for bank in s.dram]
  port = lambda name, bank: 'bank_%d_%s' % (bank, name)
  lines.append('extern "C" {')
  lines.append('')
  lines.append('void %s_kernel(\n  %s,\n  uint64_t coalesced_data_num)' % (
      stencil.app_name, ',\n  '.join(
          'ap_uint<%d>* %s' % (burst, port(n, b)) for n, b in outputs + inputs)))
  lines.append('{')
  for name, bank in inputs + outputs:
    lines.append('  hls::stream<Data<ap_uint<{0}>>> {1}_buf("{1}_buf");'.format(
        burst, port(name, bank)))
    lines.append('#pragma HLS stream variable = %s_buf depth = 32' %
                 port(name, bank))
  lines.append('')
  for module in modules:
    for fifo in module.outputs:
      lines.append('  hls::stream<Data<{0}>> {1}("{1}");'.format(
          fifo.ctype, fifo.name))
      lines.append('#pragma HLS stream variable = %s depth = 2' % fifo.name)
  lines.append('')
  lines.append('#pragma HLS dataflow')
  for name, bank in inputs:
    lines.append('  BurstRead({0}_buf, {0}, coalesced_data_num);'.format(
        port(name, bank)))
  for module in modules:
    args = ['/*output*/ ' + f.name for f in module.outputs]
    if module.kind == 'store':
      args.append('/*output*/ %s_buf' % port(module.var, module.bank))
    args += ['/* input*/ ' + f.name for f in module.inputs]
    if module.kind == 'load':
      args.append('/* input*/ %s_buf' % port(module.var, module.bank))
    lines.append('  Module%dFunc(  // %s\n    %s);' %
                 (module_func[module], module.name, ',\n    '.join(args)))
  for name, bank in outputs:
    lines.append('  BurstWrite({0}, {0}_buf, coalesced_data_num);'.format(
        port(name, bank)))
  lines.append('}')
  lines.append('')
  lines.append('}  // extern "C"')
  return '\n'.join(lines) + '\n'


def summary(stencil) -> Dict[str, int]:
  """Module / FIFO counts of the graph (what the reference logs as its
  dataflow graph)."""
  modules = build_graph(stencil)
  counts = collections.Counter(m.kind for m in modules)
  counts['fifos'] = sum(len(m.outputs) for m in modules)
  counts['delay_elements'] = sum(m.delay for m in modules)
  return dict(counts)


def build(stencil, timed: bool = False) -> str:
  source = emit(stencil)
  # no -march=native: the library is built where the repo is built and may
  # run on another host; the kernel is queue-bound, not vector code
  flags = ['-O3'] if timed else ['-O2']
  flags += ['-ffp-contract=off', '-fno-fast-math']
  digest = hashlib.sha1((source + ' '.join(flags)).encode()).hexdigest()[:12]
  for name in sorted(os.listdir(SHIM_DIR)):
    with open(os.path.join(SHIM_DIR, name), 'rb') as fp:
      digest = hashlib.sha1(digest.encode() + fp.read()).hexdigest()[:12]
  os.makedirs(BUILD_DIR, exist_ok=True)
  base = os.path.join(BUILD_DIR, 'dataflow_%s_%s' % (stencil.app_name, digest))
  lib = base + '.so'
  if os.path.exists(lib):
    return lib
  with open(base + '.cpp', 'w') as fp:
    fp.write(source)
  tmp = '%s.%d.tmp.so' % (base, os.getpid())
  subprocess.run(['g++', '-std=c++17', '-shared', '-fPIC', '-I', SHIM_DIR] +
                 flags + [base + '.cpp', '-o', tmp], check=True)
  os.replace(tmp, lib)
  return lib


class DataflowKernel:
  """ctypes wrapper with the interface of oracle.stream_kernel.StreamKernel:
  bank buffers in, bank buffers out."""

  def __init__(self, stencil, timed: bool = False):
    self.stencil = stencil
    self.lib = ctypes.CDLL(build(stencil, timed))
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
