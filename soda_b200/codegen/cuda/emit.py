"""Specialises the hand-written CUDA templates for one SODA program.

The generated translation unit contains *only*

* one functor per statement: the statement's expression printed from the IR,
  with every tensor load spelled ``a.template ld<slot, dx, dy, ds>()``;
* constexpr plan tables per pass variant (``soda::NodeDesc`` rows, halos, lags);
* the static program description consumed by the runtime (names, dtypes, valid
  boxes, pass schedule) and the program-named C entry point.

Tiling, pipelines, sliding windows, shuffles, stores and the host runtime are
the hand-written headers under ``soda_b200/csrc``; nothing kernel-shaped is
emitted here (BASELINE.json north_star: "specialised from the IR by
instantiating a hand-written template rather than by emitting free-form
kernels").  The reference's counterpart is the HLS kernel printer
(reference: src/soda/codegen/xilinx/hls_kernel.py:338-971), which does print
free-form modules.
"""
import logging
import math
from typing import Dict, List, Optional, Sequence

from soda_b200 import ir, util
from soda_b200.codegen.cuda import plan as planner
from soda_b200.optimization import fixed_point, widths

DTYPE_CODES = {
    'uint8': 'SODA_CUDA_U8',
    'int8': 'SODA_CUDA_I8',
    'uint16': 'SODA_CUDA_U16',
    'int16': 'SODA_CUDA_I16',
    'uint32': 'SODA_CUDA_U32',
    'int32': 'SODA_CUDA_I32',
    'uint64': 'SODA_CUDA_U64',
    'int64': 'SODA_CUDA_I64',
    'half': 'SODA_CUDA_F16',
    'float': 'SODA_CUDA_F32',
    'double': 'SODA_CUDA_F64',
}

_HELPERS = r'''
namespace soda_gen {
template <typename T> __host__ __device__ __forceinline__ T soda_min(T a, T b) { return b < a ? b : a; }
template <typename T> __host__ __device__ __forceinline__ T soda_max(T a, T b) { return a < b ? b : a; }
template <typename T> __host__ __device__ __forceinline__ T soda_abs(T a) { return a < 0 ? T(-a) : a; }
// integer division by zero can only happen on cells that are never stored
template <typename T> __host__ __device__ __forceinline__ T soda_div(T a, T b) { return b == 0 ? T(0) : T(a / b); }
template <typename T> __host__ __device__ __forceinline__ T soda_mod(T a, T b) { return b == 0 ? T(0) : T(a % b); }
}  // namespace soda_gen
'''


class _FunctorPrinter(ir.CPrinter):
  """CPrinter that guards integer ``/`` and ``%`` against zero divisors (they
  occur only in never-stored halo cells) - C++ semantics are unchanged for
  non-zero divisors."""

  pow2_fma = False  # --cuda-pow2-fma: see soda::fma_pow2 (soda_stream.cuh)

  @staticmethod
  def _pow2_product(node):
    """(coefficient literal, other factor) of ``c * x`` / ``x * c`` with ``c`` a
    float literal that is a power of two and ``x`` a float value, else None."""
    node = ir.unwrap(node)
    if not isinstance(node, ir.MulDiv) or tuple(node.operator) != ('*',) or \
        len(node.operand) != 2 or ir.result_type(node) != ir.FLOAT:
      return None
    for lit, other in (node.operand, node.operand[::-1]):
      lit = ir.unwrap(lit)
      if isinstance(lit, ir.Num) and lit.literal_type == ir.FLOAT and \
          ir.result_type(other) == ir.FLOAT:
        value = float(lit.value)
        if value > 0 and math.frexp(value)[0] == 0.5:
          return lit, other
    return None

  def __call__(self, node):
    if self.pow2_fma and isinstance(node, ir.AddSub) and not node.singleton \
        and ir.result_type(node) == ir.FLOAT:
      # left to right, as C++ evaluates the chain; a term `c * x` with a
      # power-of-two literal c joins the running sum in one fused operation
      text = self(node.operand[0])
      for operator, operand in zip(node.operator, node.operand[1:]):
        product = self._pow2_product(operand)
        if product is not None and ir.result_type(operand) == ir.FLOAT:
          lit, other = product
          text = 'soda::fma_pow2({}{}, {}, {})'.format(
              '-' if operator == '-' else '', lit.c_literal, self(other), text)
        else:
          text = '({} {} {})'.format(text, operator, self(operand))
      return text
    if isinstance(node, ir.Cast):
      # spelled so that the same functor text works for scalar cells and for
      # packed fp32 pairs (soda::cast_to<float>(F2) is the identity)
      return 'soda::cast_to<{}>({})'.format(node.haoda_type.c_type,
                                            self(node.expr))
    if isinstance(node, ir.MulDiv) and not node.singleton and \
        not ir.result_type(node).is_float and \
        any(op in ('/', '%') for op in node.operator):
      text = self(node.operand[0])
      t = ir.result_type(node.operand[0])
      for operator, operand in zip(node.operator, node.operand[1:]):
        t = ir.common_type(t, ir.result_type(operand))
        rhs = self(operand)
        if operator == '*':
          text = '({} * {})'.format(text, rhs)
        else:
          func = 'soda_gen::soda_div' if operator == '/' else \
              'soda_gen::soda_mod'
          text = '{}<{}>({}, {})'.format(func, t.c_type, text, rhs)
      return text
    return super().__call__(node)


def _functor(desc: planner.StageDesc, dim: int,
             param_sizes: Optional[Dict[str, List[int]]] = None,
             pow2_fma: bool = False) -> List[str]:
  stmt = desc.stmt
  slots = {name: k for k, name in enumerate(desc.slots)}
  st_idx = stmt.ref.idx
  param_sizes = param_sizes or {}

  def ref_printer(ref: ir.Ref) -> str:
    if ref.name in param_sizes:
      # p(i, j) is the constant p[i][j]
      # (reference: src/soda/codegen/frt/host.py:580-586)
      sizes = param_sizes[ref.name]
      if len(ref.idx) != len(sizes) or any(
          not 0 <= i < n for i, n in zip(ref.idx, sizes)):
        raise util.SemanticError('param reference %s is out of range' % ref)
      linear = 0
      for index, size in zip(ref.idx, sizes):
        linear = linear * size + index
      return 'soda_gen::param_%s[%d]' % (ref.name, linear)
    delta = [a - b for a, b in zip(ref.idx, st_idx)]
    dx = delta[0]
    dy = delta[1] if dim == 3 else 0
    ds = delta[dim - 1]
    return 'a.template ld<{}, {}, {}, {}>()'.format(slots[ref.name], dx, dy, ds)

  printer = _FunctorPrinter(ref_printer,
                            min_name='soda_gen::soda_min',
                            max_name='soda_gen::soda_max')
  printer.pow2_fma = pow2_fma
  ctype = stmt.haoda_type.c_type
  lines = [
      '// {}'.format(str(stmt).replace('\n', '\n// ')),
      'template <> struct Stage<{}> {{'.format(desc.index),
      '  template <class A>',
      '  static __device__ __forceinline__ auto eval(const A& a) {',
  ]
  for let in stmt.let:
    lines.append('    const auto {} = soda::cast_to<{}>({});'.format(
        let.name, let.haoda_type.c_type, printer(let.expr)))
  lines.append('    return soda::cast_to<{}>({});'.format(
      ctype, printer(stmt.expr)))
  lines.append('  }')
  lines.append('};')
  return lines


def _lcm(values: Sequence[int]) -> int:
  result = 1
  for v in values:
    result = result * v // math.gcd(result, v)
  return result


def tuning_2d(pass_plan: planner.PassPlan, options: Dict) -> Dict[str, int]:
  """Launch-shape constants of the 2-D template."""
  rings = sorted({n.ring for n in pass_plan.nodes})
  taps = sum(sum(len(d) for d in n.deltas) for n in pass_plan.nodes)
  period = _lcm([r for r in rings if r <= 4]) or 1
  chunk = options.get('chunk') or 6
  # a chunk that is a multiple of the window depths lets the unrolled step
  # loop rotate the register windows by renaming
  chunk = max(period, (chunk + period - 1) // period * period)
  chunk = min(chunk, 256)
  return {
      # measured on B200 (jacobi2d, time block 6): with 8-cell lanes two-warp
      # CTAs are 3 % faster than four-warp ones (more CTAs per SM to balance)
      'kWarps': options.get('warps') or (2 if pass_plan.cells * max(
          n.haoda_type.width_in_bits for n in pass_plan.nodes) > 128 else 4),
      'kCy': 1,
      'kMinBlocks': options.get('min_blocks') or 1,
      # 4 slots of 16-byte lanes, 3 of 32-byte lanes: ~50-70 KB per CTA
      'kStages': options.get('stages') or (4 if pass_plan.cells * max(
          n.haoda_type.width_in_bits for n in pass_plan.nodes) <= 128 else 3),
      'kChunk': chunk,
      'kUnroll': chunk,
      # the rows of a chunk are unrolled (window rotation by renaming) unless
      # the pass is huge (contrast: 197 taps - unrolled it spills 12 KB per
      # thread; erosion / xcorr with 19-deep windows but 19 taps are 1.5x
      # faster unrolled)
      'kRowUnroll': options.get('row_unroll') or (1 if taps > 100 else chunk),
  }


def _min_blocks_3d(pass_plan: planner.PassPlan) -> int:
  """CTAs per SM to ask the compiler for: small CTAs are cheap to keep
  resident; the estimate of the register windows keeps it from spilling."""
  threads = pass_plan.rows // pass_plan.cy * 32
  blocks = 4 if threads <= 128 else 2 if threads <= 256 else 1
  window = sum(n.ring * max(1, n.haoda_type.width_in_bits // 32)
               for n in pass_plan.nodes) * pass_plan.cells * pass_plan.cy
  while blocks > 1 and window + 40 > 65536 // (blocks * threads):
    blocks -= 1
  return blocks


def tuning_3d(pass_plan: planner.PassPlan, options: Dict) -> Dict[str, int]:
  geometry = planner.smem_geometry_3d(pass_plan, options.get('lookahead') or 2)
  in_depth = geometry['in_depth']
  guard = geometry['guard']
  stages = geometry['stages']
  threads = pass_plan.rows // pass_plan.cy * 32
  if threads > planner.MAX_CTA_THREADS:
    raise util.SemanticError(
        'a tile of %d rows with %d rows per thread needs %d threads per CTA '
        '(at most %d): lower --cuda-tile rows or raise the rows per thread' %
        (pass_plan.rows, pass_plan.cy, threads, planner.MAX_CTA_THREADS))
  if geometry['bytes'] > planner.SMEM_LIMIT_BYTES:
    raise util.SemanticError(
        'a tile of %d rows needs %d bytes of shared memory per CTA (at most '
        '%d): lower the tile rows or the time block' %
        (pass_plan.rows, geometry['bytes'], planner.SMEM_LIMIT_BYTES))
  # unroll the step loop so that window rotation is register renaming and ring
  # slots are compile-time constants, unless the DAG is large (code size)
  rings = sorted({n.ring for n in pass_plan.nodes if n.ring > 1})
  depths = sorted(
      set(rings) | {stages} |
      {n.smem_depth for n in pass_plan.nodes
       if n.kind == 'stage' and n.smem_depth > 0})
  unroll = options.get('unroll')
  if not unroll:
    work = sum(1 for n in pass_plan.nodes if n.kind == 'stage') * pass_plan.cy
    unroll = 1
    # ring depths first (rotation by renaming); shared-memory slot counts
    # too if that stays small (compile-time slots), else they are computed
    for candidate in (_lcm(depths), _lcm(rings) if rings else 1):
      if candidate <= 6 and candidate * work <= 64:
        unroll = candidate
        break
  return {
      'kRows': pass_plan.rows,
      'kCy': pass_plan.cy,
      'kWarps': pass_plan.rows // pass_plan.cy,
      'kUnroll': unroll,
      'kMinBlocks': options.get('min_blocks') or _min_blocks_3d(pass_plan),
      'kStages': stages,
      'kInDepth': in_depth,
      'kGuardBytes': guard,
  }


def _emit_pass(ns: str, stencil, pass_plan: planner.PassPlan,
               options: Dict) -> List[str]:
  dim = pass_plan.dim
  nodes = pass_plan.nodes
  lines = ['namespace {} {{'.format(ns)]
  lines.append('template <int N> struct NodeType;')
  for node in nodes:
    lines.append('template <> struct NodeType<{}> {{ using type = {}; }};  // {}'
                 .format(node.id, node.haoda_type.c_type, node.name))
  tuning = tuning_2d(pass_plan, options) if dim == 2 else tuning_3d(
      pass_plan, options)
  lines.append('struct Prog {')
  consts = {
      'kDim': dim,
      'kCells': pass_plan.cells,
      'kStrip': pass_plan.strip,
      'kTimeBlock': pass_plan.time_block,
      'kNumInputs': pass_plan.num_inputs,
      'kNumOutputs': pass_plan.num_outputs,
      'kNumNodes': len(nodes),
      'kHaloLo0': pass_plan.halo_lo[0],
      'kValid0': pass_plan.valid[0],
      'kAlign0': pass_plan.align0,
      'kPack': pass_plan.pack,
      'kSkew': pass_plan.skew,
      'kLoS': pass_plan.lo_s,
      'kMaxLag': pass_plan.max_lag,
  }
  if dim == 3:
    consts['kHaloLo1'] = pass_plan.halo_lo[1]
    consts['kValid1'] = pass_plan.valid[1]
  consts.update(tuning)
  for key, value in consts.items():
    lines.append('  static constexpr int {} = {};'.format(key, value))
  lines.append('  template <int N> using T = typename NodeType<N>::type;')
  lines.append('  template <int F> using StageF = soda_gen::Stage<F>;')
  lines.append('  static constexpr soda::NodeDesc kNodes[kNumNodes] = {')
  for node in nodes:
    if len(node.prods) > 8:
      raise util.SemanticError(
          'statement %s loads more than 8 distinct tensors' % node.name)
    prods = ', '.join(map(str, node.prods + [0] * (8 - len(node.prods))))
    lines.append(
        '      {{{kind}, {src}, {lag}, {ring}, {out}, {smem}, {reach}, {nprod}, {{{prods}}}}},'
        '  // {id}: {name}'.format(kind=0 if node.kind == 'input' else 1,
                                  src=node.src,
                                  lag=node.lag,
                                  ring=node.ring,
                                  out=node.out,
                                  smem=node.smem_depth,
                                  reach=node.smem_reach,
                                  nprod=len(node.prods),
                                  prods=prods,
                                  id=node.id,
                                  name=node.name))
  lines.append('  };')
  out_nodes = sorted(pass_plan.output_nodes, key=lambda n: n.out)
  lines.append('  static constexpr int kOutputNode[kNumOutputs] = {%s};' %
               ', '.join(str(n.id) for n in out_nodes))
  # the ABI promises reach_lo <= 0 <= reach_hi (include/soda_cuda.h): a window
  # that lies strictly on one side of the stored cell still spans the cell
  # itself as far as chunking and halo exchange are concerned
  reach_lo = [min(0, min(n.win_lo[d] for n in out_nodes)) for d in range(dim)]
  reach_hi = [max(0, max(n.win_hi[d] for n in out_nodes)) for d in range(dim)]
  lines.append('  static constexpr int kReachLo[%d] = {%s};' %
               (dim, ', '.join(map(str, reach_lo))))
  lines.append('  static constexpr int kReachHi[%d] = {%s};' %
               (dim, ', '.join(map(str, reach_hi))))
  lines.append('};')
  if dim == 3:
    # the planner sized the tile with its own copy of this arithmetic
    lines.append('static_assert(soda::Smem3D<Prog>::kBarrierOffset == %d, '
                 '"planner and template disagree on the shared-memory '
                 'footprint");' % planner.smem_geometry_3d(
                     pass_plan, options.get('lookahead') or 2)['barrier_offset'])
  lines.append('}  // namespace %s' % ns)
  return lines


def _c_string(text: str) -> str:
  return '\n'.join('    "%s\\n"' % line.replace('\\', '\\\\').replace('"', '\\"')
                   for line in text.split('\n'))


def emit_program(stencil,
                 time_block: Optional[int] = None,
                 options: Optional[Dict] = None) -> str:
  """Returns the text of the generated .cu file for ``stencil``."""
  options = dict(options or {})
  # fixed-point types become scaled integers, integer widths C++ does not
  # have become containers + explicit wraps
  stencil = widths.lower(fixed_point.lower(stencil))
  source_dim = stencil.dim
  if stencil.dim == 1:
    # a 1-D program runs as the 2-D program over an N x 1 grid
    from soda_b200.optimization import lift
    stencil = lift.lift_1d(stencil)
  dim = stencil.dim
  for what, stmts in (('inputs', stencil.input_stmts),
                      ('outputs', stencil.output_stmts),
                      ('params', stencil.param_stmts)):
    if len(stmts) > 8:  # SODA_CUDA_MAX_TENSORS: fixed-size tables of the ABI
      raise util.SemanticError('the CUDA backend supports at most 8 %s, the '
                               'program has %d' % (what, len(stmts)))
  if getattr(stencil, 'preserve_border', False):
    # "Reserved" in the reference (src/soda/core.py:30); its FRT host - the path
    # this backend replaces - ignores it too (only the legacy Xilinx host reads
    # it, src/soda/codegen/xilinx/host.py:855).  Say so instead of silence.
    logging.getLogger(__name__).warning(
        'border: preserve has no effect in the CUDA backend: cells outside '
        'the valid box keep the caller\'s bytes (as with border: ignore)')
  time_block = planner.choose_time_block(stencil, time_block, options)
  schedule = planner.pass_schedule(stencil.iterate, time_block)
  variants = sorted(set(schedule), reverse=True)
  plans = {
      tb: planner.make_tuned_pass_plan(stencil, tb, options) for tb in variants
  }
  stages = plans[variants[0]].stages

  lines = [
      '// Generated by soda_b200.codegen.cuda from the SODA program below.',
      '// Only functors and constexpr plan tables are generated; the kernels are',
      '// the hand-written templates in soda_stream.cuh.',
      '/*',
      str(stencil),
      '*/',
      '#include <math.h>',
      '#include "soda_stream.cuh"',
      _HELPERS,
  ]
  # `param` arrays live in constant memory: every lane reads the same element
  param_sizes = {s.name: [int(x) for x in s.size] for s in stencil.param_stmts}
  if param_sizes:
    lines.append('namespace soda_gen {')
    for stmt in stencil.param_stmts:
      elems = 1
      for n in param_sizes[stmt.name]:
        elems *= n
      lines.append('// %s' % stmt)
      lines.append('SODA_CONSTANT %s param_%s[%d];' %
                   (stmt.haoda_type.c_type, stmt.name, elems))
    lines.append('}  // namespace soda_gen')
  lines += [
      'namespace soda_gen {',
      'template <int F> struct Stage;',
  ]
  for desc in stages:
    lines.extend(_functor(desc, dim, param_sizes,
                          pow2_fma=bool(options.get('pow2_fma'))))
  lines.append('}  // namespace soda_gen')
  lines.append('')
  lines.append('namespace soda_gen {')
  for tb in variants:
    lines.extend(_emit_pass('tb%d' % tb, stencil, plans[tb], options))
  lines.append('}  // namespace soda_gen')
  lines.append('')

  # ---- host side -----------------------------------------------------------
  lines.append('#include "soda_runtime.cuh"')
  lines.append('')
  lines.append('static const soda::rt::ProgramDesc& soda_program() {')
  lines.append('  static const soda::rt::PassImpl impls[] = {')
  for tb in variants:
    lines.append('      soda::rt::make_pass_impl<soda_gen::tb%d::Prog>(),' % tb)
  lines.append('  };')
  lines.append('  static const int schedule[] = {%s};' %
               ', '.join(str(variants.index(tb)) for tb in schedule))
  lines.append('  static const soda::rt::ProgramDesc desc = [] {')
  lines.append('    soda::rt::ProgramDesc d;')
  lines.append('    memset(&d, 0, sizeof(d));')
  lines.append('    d.info.app_name = "%s";' % stencil.app_name)
  lines.append('    d.info.soda_source =\n%s;' % _c_string(str(stencil)))
  lines.append('    d.info.dim = %d;' % dim)
  lines.append('    d.info.iterate = %d;' % stencil.iterate)
  lines.append('    d.info.num_inputs = %d;' % len(stencil.input_stmts))
  lines.append('    d.info.num_outputs = %d;' % len(stencil.output_stmts))
  bytes_per_cell = 0
  for i, stmt in enumerate(stencil.input_stmts):
    lines.append('    d.info.input_names[%d] = "%s";' % (i, stmt.name))
    lines.append('    d.info.input_dtypes[%d] = %s;' %
                 (i, DTYPE_CODES[str(stmt.haoda_type)]))
    lines.append('    d.in_elem_bytes[%d] = %d;' %
                 (i, stmt.haoda_type.width_in_bits // 8))
    bytes_per_cell += stmt.haoda_type.width_in_bits // 8
  for o, stmt in enumerate(stencil.output_stmts):
    lines.append('    d.info.output_names[%d] = "%s";' % (o, stmt.name))
    lines.append('    d.info.output_dtypes[%d] = %s;' %
                 (o, DTYPE_CODES[str(stmt.haoda_type)]))
    lines.append('    d.out_elem_bytes[%d] = %d;' %
                 (o, stmt.haoda_type.width_in_bits // 8))
    bytes_per_cell += stmt.haoda_type.width_in_bits // 8
    lo, hi = stencil.window_bounds[stmt.name]
    for d in range(dim):
      lines.append('    d.info.final_lo[%d][%d] = %d;' % (o, d, max(0, -lo[d])))
      lines.append('    d.info.final_hi[%d][%d] = %d;' % (o, d, max(0, hi[d])))
  lines.append('    d.info.num_passes = %d;' % len(schedule))
  lines.append('    d.info.strict_fp = %d;' %
               (0 if options.get('fast_fp') else 1))
  lines.append('    d.info.algorithmic_bytes_per_cell_per_pass = %d;' %
               bytes_per_cell)
  lines.append('    d.info.source_dim = %d;' % source_dim)
  lines.append('    d.info.num_params = %d;' % len(stencil.param_stmts))
  for k, stmt in enumerate(stencil.param_stmts):
    elems = 1
    for n in param_sizes[stmt.name]:
      elems *= n
    lines.append('    d.info.param_names[%d] = "%s";' % (k, stmt.name))
    lines.append('    d.info.param_dtypes[%d] = %s;' %
                 (k, DTYPE_CODES[str(stmt.haoda_type)]))
    lines.append('    d.info.param_elems[%d] = %d;' % (k, elems))
    lines.append('    d.param_symbol[%d] = &soda_gen::param_%s;' %
                 (k, stmt.name))
    lines.append('    d.param_bytes[%d] = %d;' %
                 (k, elems * (stmt.haoda_type.width_in_bits // 8)))
  lines.append('    d.impls = impls;')
  lines.append('    d.num_impls = %d;' % len(variants))
  lines.append('    d.schedule = schedule;')
  lines.append('    return d;')
  lines.append('  }();')
  lines.append('  return desc;')
  lines.append('}')
  lines.append('')

  # program-named entry point mirroring soda::app::<app>
  params = []
  for stmt in stencil.input_stmts:
    params.append('const %s* var_%s_ptr' % (stmt.haoda_type.c_type, stmt.name))
    params.extend('const int32_t* var_%s_%s' % (stmt.name, what)
                  for what in ('extent', 'stride', 'min'))
  for stmt in stencil.output_stmts:
    params.append('%s* var_%s_ptr' % (stmt.haoda_type.c_type, stmt.name))
    params.extend('const int32_t* var_%s_%s' % (stmt.name, what)
                  for what in ('extent', 'stride', 'min'))
  for stmt in stencil.param_stmts:  # after the tensors, src/.../host.py:73
    params.append('const %s* var_%s_ptr' % (stmt.haoda_type.c_type, stmt.name))
    params.extend('const int32_t* var_%s_%s' % (stmt.name, what)
                  for what in ('extent', 'stride', 'min'))
  params.append('const soda_cuda_opts* opts')
  lines.append('// reference: soda::app::%s (src/soda/codegen/frt/host.py:62-89)' %
               stencil.app_name)
  lines.append('extern "C" SODA_CUDA_API int soda_cuda_%s(\n    %s) {' %
               (stencil.app_name, ',\n    '.join(params)))
  first = stencil.input_stmts[0].name
  for stmt in stencil.input_stmts + stencil.output_stmts:
    for what in ('min',):
      lines.append('  (void)var_%s_%s;' % (stmt.name, what))
  lines.append('  if (var_%s_extent == nullptr)' % first)
  lines.append('    return soda::rt::fail(SODA_CUDA_BAD_ARGUMENT, '
               '"extent is NULL");')
  if source_dim == 1:
    # the caller's quadruples are 1-D (the program as written): N cells,
    # stride 1; the library runs the N x 1 grid
    lines.append('  const int32_t soda_extent[2] = {var_%s_extent[0], 1};' % first)
    lines.append('  const int32_t soda_stride[2] = {1, var_%s_extent[0]};' % first)
    for stmt in stencil.input_stmts + stencil.output_stmts:
      lines.append('  if (var_%s_stride != nullptr && var_%s_stride[0] != 1)' %
                   (stmt.name, stmt.name))
      lines.append('    return soda::rt::fail(SODA_CUDA_UNSUPPORTED, '
                   '"1-D arrays must be dense");')
  for stmt in stencil.input_stmts[1:] + stencil.output_stmts:
    lines.append('  if (var_%s_extent != nullptr)' % stmt.name)
    lines.append('    for (int d = 0; d < %d; ++d)' % source_dim)
    lines.append('      if (var_%s_extent[d] != var_%s_extent[d])' %
                 (stmt.name, first))
    lines.append('        return soda::rt::fail(SODA_CUDA_BAD_ARGUMENT, '
                 '"all tensors must share one extent");')
  for k, stmt in enumerate(stencil.param_stmts):
    # params are small dense arrays (the reference declares them as C arrays,
    # src/soda/codegen/frt/host.py:497-500); extent / stride / min are accepted
    # for signature compatibility
    for what in ('extent', 'stride', 'min'):
      lines.append('  (void)var_%s_%s;' % (stmt.name, what))
    lines.append('  { int status = soda_cuda_set_param(%d, var_%s_ptr, opts); '
                 'if (status != SODA_CUDA_OK) return status; }' % (k, stmt.name))
  lines.append('  const void* in_ptrs[] = {%s};' %
               ', '.join('var_%s_ptr' % s.name for s in stencil.input_stmts))
  stride_of = (lambda s: 'soda_stride') if source_dim == 1 else (
      lambda s: 'var_%s_stride' % s.name)
  lines.append('  const int32_t* in_strides[] = {%s};' %
               ', '.join(stride_of(s) for s in stencil.input_stmts))
  lines.append('  void* out_ptrs[] = {%s};' %
               ', '.join('var_%s_ptr' % s.name for s in stencil.output_stmts))
  lines.append('  const int32_t* out_strides[] = {%s};' % ', '.join(
      stride_of(s) for s in stencil.output_stmts))
  gpus = int(options.get('gpus') or 0)
  if gpus > 1:
    # sodac --cuda-gpus N: the default of this library, unless the caller's
    # opts name a device count themselves
    lines.append('  soda_cuda_opts soda_opts;')
    lines.append('  memset(&soda_opts, 0, sizeof(soda_opts));')
    lines.append('  soda_opts.device = -1;')
    lines.append('  if (opts != nullptr) soda_opts = *opts;')
    lines.append('  soda_opts.struct_size = sizeof(soda_opts);')
    lines.append('  if (soda_opts.reserved[1] == 0) soda_opts.reserved[1] = %d;'
                 % gpus)
    lines.append('  opts = &soda_opts;')
  lines.append('  return soda_cuda_run_host(in_ptrs, in_strides, out_ptrs, '
               'out_strides, %s, opts);' %
               ('soda_extent' if source_dim == 1 else 'var_%s_extent' % first))
  lines.append('}')
  return '\n'.join(lines) + '\n'
