"""ORACLE (test infrastructure, not product): NumPy restatement of the golden
loops that the reference's generated test ``main`` runs.

PARITY PINNING STATUS: see oracle/README.md.  The reference cannot be built or
imported in this image (its un-vendored ``haoda``/``textx``/``pulp``
dependencies and the Xilinx headers are absent), so this file is pinned
against (a) hand-computed vectors in tests/golden/, (b) the loop bounds and
load addresses printed by the reference's own ``print_test`` code run with
stand-in printers (oracle/pin_against_reference.py), and (c) the independent
g++-compiled restatement in oracle/emit_cpp.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module.  The product path (soda_b200.codegen.cuda) never does.

Semantics restated (reference: src/soda/codegen/frt/host.py:556-624):

  for each non-input tensor X in chronological order:
    for p in valid_box(X):            # reference :565-578
      X[p] = T_X( expr( parent[p + (ld_idx - st_idx)] ) )     # :587-592,616-623

where ``valid_box`` is ``[max(0,-lo_d), extent_d - max(0,hi_d))`` with lo/hi the
bounds of the overall window inputs->X, every tensor is a dense array of the
grid's extent, and cells outside a tensor's box are never written.  Arithmetic
follows the C++ usual arithmetic conversions (what g++ computes for the text the
reference prints); float ops are IEEE, un-contracted, in IR order.

Arrays are NumPy C-order with shape ``extent[::-1]`` (dimension 0 contiguous).
"""
from typing import Dict, Sequence

import numpy as np

from soda_b200 import ir

_NP = {
    'float': np.float32,
    'double': np.float64,
    # IEEE binary16: NumPy rounds every float16 operation once, to nearest-even
    'half': np.float16,
    'bool': np.bool_,
}


def np_dtype(t: ir.Type):
  """NumPy type that holds a value of ``t`` (for ``uintN`` / ``intN`` that C++
  does not have: the container the reference's ``ap_uint<N>`` array occupies)."""
  name = str(t)
  if name in _NP:
    return _NP[name]
  return np.dtype(t.container.numpy_name).type


def _wrap(arr, t: ir.Type):
  """Keeps the low N bits of an integer array, sign-extended for ``intN``: what
  assigning to ``ap_uint<N>`` / ``ap_int<N>`` does."""
  bits = t.width_in_bits
  wide = np.asarray(arr).astype(np.int64)
  wide = wide & ((1 << bits) - 1)
  if t.is_signed:
    sign = 1 << (bits - 1)
    wide = (wide ^ sign) - sign
  return wide.astype(np_dtype(t))


def _as(value, t: ir.Type):
  """C++ conversion of ``value`` (array or scalar) to type ``t``."""
  dtype = np_dtype(t)
  arr = np.asarray(value)
  if t.is_lowerable:
    if arr.dtype.kind == 'f':
      with np.errstate(invalid='ignore'):
        arr = np.trunc(arr).astype(np.int64)
    return _wrap(arr, t)
  if arr.dtype == dtype:
    return arr
  if arr.dtype.kind == 'f' and np.dtype(dtype).kind in 'iu':
    # float -> int truncates toward zero (UB when out of range; tests stay in
    # range)
    with np.errstate(invalid='ignore'):
      return np.trunc(arr).astype(np.int64).astype(dtype)
  if np.dtype(dtype).kind == 'b':
    return arr != 0
  with np.errstate(over='ignore', invalid='ignore'):
    return arr.astype(dtype)


def _raw(value, t: ir.Type):
  """(scaled int64 integers, fractional bits) of a fixed-point or integer
  value."""
  if t is not None and t.is_float:
    raise TypeError('fixed-point arithmetic mixed with a floating-point value; '
                    'cast one side explicitly')
  return np.asarray(value).astype(np.int64), (t.frac_bits if t is not None
                                               else 0)


def _convert(value, source: ir.Type, t: ir.Type):
  """Conversion of ``value`` of type ``source`` to ``t`` where one of them is a
  fixed-point type (``ap_fixed`` defaults: AP_TRN, AP_WRAP; restated from the
  public documentation, see soda_b200/optimization/fixed_point.py - which this
  evaluator deliberately does not use)."""
  if t.is_fixed:
    frac = t.frac_bits
    if source is not None and source.is_float:
      with np.errstate(invalid='ignore'):
        raw = np.floor(np.asarray(value).astype(np.float64) *
                       float(1 << frac)).astype(np.int64)
    else:
      raw, have = _raw(value, source)
      if have >= frac:
        raw = raw >> (have - frac)  # arithmetic shift: toward minus infinity
      else:
        raw = raw << (frac - have)
    return _wrap(raw, t.raw_type)
  raw, have = _raw(value, source)
  if t.is_float:
    return (raw.astype(np.float64) / float(1 << have)).astype(np_dtype(t))
  # to an integer: toward zero
  magnitude = np.abs(raw) >> have
  return _as(np.where(raw < 0, -magnitude, magnitude), t)


def _to(value, source: ir.Type, t: ir.Type):
  if t.is_fixed or (source is not None and source.is_fixed):
    return _convert(value, source, t)
  return _as(value, t)


def _trunc_div(a, b):
  """C++ integer division (toward zero); x/0 yields 0 instead of trapping."""
  if a.dtype.kind == 'u':
    safe = np.where(b == 0, 1, b).astype(a.dtype)
    return np.where(b == 0, 0, a // safe).astype(a.dtype)
  safe = np.where(b == 0, 1, b)
  q = np.abs(a.astype(np.int64)) // np.abs(safe.astype(np.int64))
  q = np.where((a < 0) != (safe < 0), -q, q)
  return np.where(b == 0, 0, q).astype(a.dtype)


def _trunc_mod(a, b):
  safe = np.where(b == 0, 1, b).astype(a.dtype)
  return np.where(b == 0, 0, np.fmod(a, safe)).astype(a.dtype)


class _Evaluator:

  def __init__(self, load, variables):
    self.load = load
    self.variables = variables

  def __call__(self, node):
    """Returns (value, Type)."""
    if isinstance(node, ir.Operand):
      return self(node.inner)
    if isinstance(node, ir.Ref):
      return self.load(node), node.haoda_type
    if isinstance(node, ir.Var):
      return self.variables[node.name]
    if isinstance(node, ir.Num):
      t = node.literal_type
      return np_dtype(t)(node.value), t
    if isinstance(node, ir.Cast):
      value, source = self(node.expr)
      return _to(value, source, node.haoda_type), node.haoda_type
    if isinstance(node, ir.Unary):
      value, t = self(node.operand)
      for op in reversed(node.operator):
        if op == '!':
          value, t = np.asarray(value) == 0, ir.Type('bool')
          continue
        if t.is_fixed:
          if op == '~':
            raise TypeError('~ of a fixed-point value')
          raw, frac = _raw(value, t)
          value, t = (-raw if op == '-' else raw), ir.Type.exact_fixed(frac)
          continue
        t = ir._promote(t)
        value = _as(value, t)
        with np.errstate(over='ignore'):
          if op == '-':
            value = np.negative(value)
          elif op == '~':
            value = np.invert(value)
      return value, t
    if isinstance(node, ir.BinaryOp):
      value, t = self(node.operand[0])
      for operator, operand in zip(node.operator, node.operand[1:]):
        rhs, rt = self(operand)
        value, t = self.binary(operator, value, t, rhs, rt)
      return value, t
    if isinstance(node, ir.Call):
      args = [self(arg) for arg in node.arg]
      if node.name in ir.DOUBLE_MATH_CALLS:
        func = {
            'fabs': np.abs,
            'pow': np.power,
        }.get(node.name) or getattr(np, node.name)
        with np.errstate(all='ignore'):
          if node.haoda_type == ir.FLOAT:  # float math mode
            return func(_as(args[0][0], ir.FLOAT)).astype(np.float32), ir.FLOAT
          return func(*[_as(v, ir.DOUBLE) for v, _ in args]), ir.DOUBLE
      if any(at.is_fixed for _, at in args):
        if node.name not in ir.SELECT_CALLS and node.name != 'abs':
          raise TypeError('%s of a fixed-point value' % node.name)
        raws = [_raw(v, at) for v, at in args]
        frac = max(f for _, f in raws)
        aligned = [r << (frac - f) for r, f in raws]
        if node.name == 'abs':
          value = np.abs(aligned[0])
        else:
          func = np.minimum if node.name == 'min' else np.maximum
          value = aligned[0]
          for other in aligned[1:]:
            value = func(value, other)
        return value, ir.Type.exact_fixed(frac)
      if node.name in ir.SELECT_CALLS:
        t = args[0][1]
        for _, at in args[1:]:
          t = ir.common_type(t, at)
        if len(args) > 1:
          t = ir._promote(t)
        func = np.minimum if node.name == 'min' else np.maximum
        value = _as(args[0][0], t)
        for v, _ in args[1:]:
          value = func(value, _as(v, t))
        return value, t
      if node.name == 'abs':
        t = ir._promote(args[0][1])
        with np.errstate(over='ignore'):
          return np.abs(_as(args[0][0], t)), t
    raise NotImplementedError('oracle cannot evaluate %r' % node)

  @staticmethod
  def binary(op, a, at, b, bt):
    if op in ('||', '&&'):
      a, b = np.asarray(a) != 0, np.asarray(b) != 0
      return (np.logical_or(a, b) if op == '||' else np.logical_and(a, b),
              ir.Type('bool'))
    if at.is_fixed or bt.is_fixed:
      (ra, fa), (rb, fb) = _raw(a, at), _raw(b, bt)
      if op == '*':
        return ra * rb, ir.Type.exact_fixed(fa + fb)
      frac = max(fa, fb)
      ra, rb = ra << (frac - fa), rb << (frac - fb)
      if op == '+':
        return ra + rb, ir.Type.exact_fixed(frac)
      if op == '-':
        return ra - rb, ir.Type.exact_fixed(frac)
      compare = {'==': np.equal, '!=': np.not_equal, '<=': np.less_equal,
                 '>=': np.greater_equal, '<': np.less, '>': np.greater}
      if op in compare:
        return compare[op](ra, rb), ir.Type('bool')
      raise TypeError('operator %s on fixed-point values' % op)
    t = ir.common_type(at, bt)
    a, b = _as(a, t), _as(b, t)
    if op in ('==', '!=', '<=', '>=', '<', '>'):
      func = {
          '==': np.equal,
          '!=': np.not_equal,
          '<=': np.less_equal,
          '>=': np.greater_equal,
          '<': np.less,
          '>': np.greater,
      }[op]
      return func(a, b), ir.Type('bool')
    with np.errstate(all='ignore'):
      if op == '+':
        return np.add(a, b), t
      if op == '-':
        return np.subtract(a, b), t
      if op == '*':
        return np.multiply(a, b), t
      if op == '/':
        if t.is_float:
          return np.divide(a, b), t
        return _trunc_div(np.asarray(a), np.asarray(b)), t
      if op == '%':
        return _trunc_mod(np.asarray(a), np.asarray(b)), t
      if op == '|':
        return np.bitwise_or(a, b), t
      if op == '^':
        return np.bitwise_xor(a, b), t
      if op == '&':
        return np.bitwise_and(a, b), t
    raise NotImplementedError(op)


def run(stencil, inputs: Dict[str, np.ndarray],
        keep_intermediates: bool = False,
        params: Dict[str, np.ndarray] = None) -> Dict[str, np.ndarray]:
  """Evaluates the whole ``iterate``-unrolled chain on ``inputs``.

  ``params``: {param name: array of the declared size}; a reference ``p(i, j)``
  is the constant ``p[i][j]`` (reference:
  src/soda/codegen/frt/host.py:580-586).

  Returns {output name: array}; cells outside an output's valid box are zero
  (the reference leaves them untouched; compare only inside
  ``stencil.valid_box``).  With ``keep_intermediates`` every tensor of the
  chain is returned.
  """
  dim = stencil.dim
  first = inputs[stencil.input_names[0]]
  extent = tuple(first.shape[::-1])
  data: Dict[str, np.ndarray] = {}
  for name, stmt in zip(stencil.input_names, stencil.input_stmts):
    arr = np.ascontiguousarray(inputs[name])
    if arr.dtype != np.dtype(np_dtype(stmt.haoda_type)):
      raise TypeError('input %s must be %s' % (name, stmt.haoda_type))
    if tuple(arr.shape[::-1]) != extent:
      raise ValueError('all inputs must share one extent')
    if stmt.haoda_type.is_lowerable:
      arr = _wrap(arr, stmt.haoda_type)  # an ap_uint<N> array holds N bits
    elif stmt.haoda_type.is_fixed and stmt.haoda_type.raw_type.is_lowerable:
      arr = _wrap(arr, stmt.haoda_type.raw_type)
    data[name] = arr

  tensors = stencil.chronological_tensors
  last_use = {}
  for i, tensor in enumerate(tensors):
    for parent in tensor.parents:
      last_use[parent] = i

  outputs = {}
  for i, tensor in enumerate(tensors):
    if tensor.is_input():
      continue
    box = stencil.valid_box(tensor.name, extent)
    out = np.zeros(extent[::-1], dtype=np_dtype(tensor.haoda_type))
    if all(hi > lo for lo, hi in box):
      st_idx = tensor.st_idx

      def load(ref, tensor=tensor, box=box, st_idx=st_idx):
        if ref.name in stencil.param_names:
          return np.asarray(params[ref.name])[tuple(ref.idx)]
        delta = tuple(a - b for a, b in zip(ref.idx, st_idx))
        index = tuple(
            slice(box[d][0] + delta[d], box[d][1] + delta[d])
            for d in reversed(range(dim)))
        return data[ref.name][index]


      variables = {}
      evaluator = _Evaluator(load, variables)
      for let in tensor.lets:
        value, t = evaluator(let.expr)
        if let.haoda_type is not None:
          value, t = _to(value, t, let.haoda_type), let.haoda_type
        variables[let.name] = (value, t)
      value, source = evaluator(tensor.expr)
      index = tuple(slice(box[d][0], box[d][1]) for d in reversed(range(dim)))
      out[index] = _to(value, source, tensor.haoda_type)
    data[tensor.name] = out
    if tensor.name in stencil.output_names or keep_intermediates:
      outputs[tensor.name] = out
    if not keep_intermediates:
      for name in list(data):
        if last_use.get(name, -1) <= i and name not in stencil.output_names \
            and name != tensor.name and name not in stencil.input_names:
          del data[name]
  return outputs


def reference_inputs(stencil, extent: Sequence[int],
                     seed: int = 0) -> Dict[str, np.ndarray]:
  """Inputs shaped like the reference test main's: integer tensors get
  ``p + q (+ r)`` (reference: src/soda/codegen/frt/host.py:519-528), float
  tensors get U[0, 1) (the reference draws from std::default_random_engine,
  which NumPy cannot reproduce bit for bit; a seeded NumPy generator is used
  instead)."""
  rng = np.random.default_rng(seed)
  result = {}
  shape = tuple(extent[::-1])
  for stmt in stencil.input_stmts:
    dtype = np_dtype(stmt.haoda_type)
    if stmt.haoda_type.is_float:
      result[stmt.name] = rng.random(shape, dtype=np.float64).astype(dtype)
    else:
      grids = np.indices(shape).sum(axis=0)
      result[stmt.name] = grids.astype(dtype)
  return result


def reference_params(stencil) -> Dict[str, np.ndarray]:
  """Params as the reference's test main fills them: ``p[x][y] = x + y``
  (reference: src/soda/codegen/frt/host.py:530-543)."""
  result = {}
  for stmt in stencil.param_stmts:
    shape = tuple(int(x) for x in stmt.size)
    result[stmt.name] = np.indices(shape).sum(axis=0).astype(
        np_dtype(stmt.haoda_type))
  return result


def default_extent(stencil):
  """The reference test main's default problem size: tile size in every tiled
  dimension and ``kStencilDim + 1`` in the last one
  (reference: src/soda/codegen/frt/host.py:453-460)."""
  lo, hi = stencil.window_bounds[stencil.output_names[0]]
  last = stencil.dim - 1
  return tuple(stencil.tile_size[:-1]) + (hi[last] - lo[last] + 1 + 1,)
