"""Fixed-point types (``uint18_3``, ``int32_16`` ...) as scaled integers.

The reference prints ``(u)intN_M`` as ``ap_(u)fixed<N, M>`` (haoda
``Type.c_type``; grammar: src/soda/grammar.py:37-45, parsed and printed back by
src/tests/test_grammar.py:40-54): N bits in all, M of them integer bits, so a
value is ``raw * 2^-(N - M)`` with ``raw`` an N-bit (two's complement when
signed) integer.  No reference test executes such a program and the arithmetic
lives in the Xilinx headers, which are not here; the rules below are the
*documented defaults* of ``ap_fixed`` (quantisation ``AP_TRN``, overflow
``AP_WRAP``) for the operations that are exact, and everything else is
rejected rather than guessed:

* ``a + b``, ``a - b``: exact, ``max(Fa, Fb)`` fractional bits;
  ``a * b``: exact, ``Fa + Fb`` fractional bits; unary ``-``; an integer
  operand has 0 fractional bits (``ap_fixed`` operators widen their result);
* comparisons align the fractional bits and compare the scaled integers;
  ``min`` / ``max`` / ``abs`` likewise;
* a store, a typed ``let`` or a cast to a fixed-point type drops surplus
  fractional bits toward minus infinity (``AP_TRN``) and keeps the low N bits,
  sign-extended for ``intN_M`` (``AP_WRAP``); from a ``float`` / ``double``
  value: ``floor(v * 2^F)``, then the same wrap;
* a cast of a fixed-point value to ``float`` / ``double`` is ``raw / 2^F``
  (computed in double), to an integer type it truncates toward zero (the C
  conversion ``ap_fixed`` follows);
* division, ``%``, bit operations and arithmetic that mixes fixed-point and
  floating-point operands without a cast raise ``SemanticError``.

This pass rewrites such a program **in the DSL itself** into one whose tensors
are the plain integers ``uintN`` / ``intN`` holding ``raw`` and whose
expressions are 64-bit integer arithmetic on them; optimization/widths.py then
gives those their containers and wraps.  Exactness contract: intermediates fit
63 bits.  The NumPy oracle (oracle/golden.py) does not use this pass: it
evaluates fixed-point programs natively, so a mistake here cannot hide; the
g++ oracle and the CUDA backend run the rewritten program.
"""
from typing import Dict, Optional, Tuple

from soda_b200 import grammar, ir, util

FIXED, INT, FLOAT, BOOL = 'fixed', 'int', 'float', 'bool'


def has_fixed_types(program_or_stencil) -> bool:
  stmts = (list(program_or_stencil.input_stmts) +
           list(program_or_stencil.param_stmts) +
           list(program_or_stencil.local_stmts) +
           list(program_or_stencil.output_stmts))
  found = []

  def look(obj, args):
    if isinstance(obj, ir.Cast) and obj.haoda_type.is_fixed:
      found.append(obj)
    return obj

  for stmt in stmts:
    if stmt.haoda_type.is_fixed:
      return True
    for let in getattr(stmt, 'let', ()):
      if let.haoda_type is not None and let.haoda_type.is_fixed:
        return True
      let.expr.visit(look)
    if getattr(stmt, 'expr', None) is not None:
      stmt.expr.visit(look)
  return bool(found)


def _kind_of(t: ir.Type) -> str:
  if t.is_fixed:
    return FIXED
  if t.is_float:
    return FLOAT
  return BOOL if t == 'bool' else INT


def _scale_up(text: str, bits: int) -> str:
  return text if bits == 0 else '(%s * %d)' % (text, 1 << bits)


class _Translator:
  """expression -> (DSL text, kind, fractional bits)."""

  def __init__(self, symbols: Dict[str, ir.Type]):
    self.symbols = symbols  # tensors and params
    self.variables: Dict[str, Tuple[str, int]] = {}  # let name -> (kind, F)
    self.lets = []  # let lines of the rewritten statement, in order

  def floor_shift(self, text: str, bits: int) -> str:
    """``floor(text / 2^bits)`` of a 64-bit integer without a shift operator
    (the DSL has none): the low bits are removed first, so ``/`` is exact.  The
    value is bound to a let so that its text appears once."""
    if bits == 0:
      return text
    name = 'fx%d' % len(self.lets)
    self.lets.append('int64 %s = %s' % (name, text))
    return '((%s - (%s & %d)) / %d)' % (name, name, (1 << bits) - 1, 1 << bits)

  def __call__(self, node) -> Tuple[str, str, int]:
    if isinstance(node, ir.Operand):
      return self(node.inner)
    if isinstance(node, ir.Ref):
      t = self.symbols[node.name]
      if t.is_fixed:
        return 'int64(%s)' % node, FIXED, t.frac_bits
      return str(node), _kind_of(t), 0
    if isinstance(node, ir.Var):
      kind, frac = self.variables.get(node.name, (None, 0))
      if kind is None:  # a param element
        t = self.symbols[node.name]
        if t.is_fixed:
          return 'int64(%s)' % node, FIXED, t.frac_bits
        return str(node), _kind_of(t), 0
      return str(node), kind, frac
    if isinstance(node, ir.Num):
      return str(node), _kind_of(node.literal_type), 0
    if isinstance(node, ir.Cast):
      text, kind, frac = self(node.expr)
      return self.convert(text, kind, frac, node.haoda_type, cast=True)
    if isinstance(node, ir.Unary):
      text, kind, frac = self(node.operand)
      for op in reversed(node.operator):
        if op == '!':
          text, kind, frac = '(!%s)' % text, BOOL, 0
        elif kind == FIXED and op == '~':
          raise util.SemanticError('~ of a fixed-point value is not supported')
        else:
          text = '(%s%s)' % (op, text)
          if kind == BOOL:
            kind = INT
      return text, kind, frac
    if isinstance(node, ir.BinaryOp):
      text, kind, frac = self(node.operand[0])
      for operator, operand in zip(node.operator, node.operand[1:]):
        rhs, rkind, rfrac = self(operand)
        text, kind, frac = self.binary(operator, text, kind, frac, rhs, rkind,
                                       rfrac)
      return text, kind, frac
    if isinstance(node, ir.Call):
      args = [self(arg) for arg in node.arg]
      if not any(kind == FIXED for _, kind, _ in args):
        kind = FLOAT if (node.name in ir.DOUBLE_MATH_CALLS or
                         any(k == FLOAT for _, k, _ in args)) else INT
        return '%s(%s)' % (node.name, ', '.join(t for t, _, _ in args)), kind, 0
      if node.name in ir.SELECT_CALLS or node.name == 'abs':
        if any(kind == FLOAT for _, kind, _ in args):
          raise util.SemanticError(
              '%s mixes fixed-point and floating-point arguments' % node.name)
        frac = max(f for _, _, f in args)
        aligned = [_scale_up('int64(%s)' % t, frac - f) for t, _, f in args]
        return '%s(%s)' % (node.name, ', '.join(aligned)), FIXED, frac
      raise util.SemanticError('%s of a fixed-point value is not supported; '
                               'cast to float first' % node.name)
    raise util.InternalError('fixed_point: cannot translate %r' % node)

  @staticmethod
  def binary(op, a, akind, afrac, b, bkind, bfrac):
    if FIXED not in (akind, bkind):
      if op in ('||', '&&', '==', '!=', '<=', '>=', '<', '>'):
        kind = BOOL
      else:
        kind = FLOAT if FLOAT in (akind, bkind) else INT
      return '(%s %s %s)' % (a, op, b), kind, 0
    if op in ('||', '&&'):
      return '((%s != 0) %s (%s != 0))' % (a, op, b), BOOL, 0
    if FLOAT in (akind, bkind):
      raise util.SemanticError(
          'operator %s mixes fixed-point and floating-point operands; cast '
          'one side explicitly' % op)
    a, b = 'int64(%s)' % a, 'int64(%s)' % b
    if op in ('+', '-', '==', '!=', '<=', '>=', '<', '>'):
      frac = max(afrac, bfrac)
      text = '(%s %s %s)' % (_scale_up(a, frac - afrac), op,
                             _scale_up(b, frac - bfrac))
      if op in ('+', '-'):
        return text, FIXED, frac
      return text, BOOL, 0
    if op == '*':
      return '(%s * %s)' % (a, b), FIXED, afrac + bfrac
    raise util.SemanticError(
        'operator %s is not supported on fixed-point values (exact operations '
        'only: + - *, comparisons, min / max / abs)' % op)

  def convert(self, text, kind, frac, t: Optional[ir.Type], cast=False):
    """Value of ``text`` converted to type ``t`` (a cast, a typed let or a
    store); the wrap to N bits is left to optimization/widths.py, which sees a
    cast / store to ``t.raw_type``."""
    if t is None:
      return text, kind, frac
    if t.is_fixed:
      if kind == FLOAT:
        raw = 'int64(floor(double(%s) * %d.0))' % (text, 1 << t.frac_bits)
      elif frac >= t.frac_bits:
        raw = self.floor_shift('int64(%s)' % text, frac - t.frac_bits)
      else:
        raw = _scale_up('int64(%s)' % text, t.frac_bits - frac)
      if cast:
        # an N-bit value again: int64(uintN(...)) so that widths.py wraps it
        raw = 'int64(%s(%s))' % (t.raw_type, raw)
      return raw, FIXED, t.frac_bits
    if kind != FIXED:
      return ('%s(%s)' % (t, text) if cast else text), _kind_of(t), 0
    if t.is_float:
      value = '(double(%s) / %d.0)' % (text, 1 << frac)
      return '%s(%s)' % (t, value), FLOAT, 0
    # fixed -> integer: toward zero, as C converts (ap_fixed::to_int)
    value = '(%s / %d)' % ('int64(%s)' % text, 1 << frac)
    return '%s(%s)' % (t, value), _kind_of(t), 0


def lower_text(text: str) -> str:
  """SODA source -> equivalent source without fixed-point types."""
  program = grammar.parse(text)
  if not has_fixed_types(program):
    return text
  symbols = {}
  for stmt in (program.input_stmts + program.param_stmts +
               program.local_stmts + program.output_stmts):
    symbols[stmt.name] = stmt.haoda_type
  out = [l for l in str(program).split('\n')
         if l.strip() and not l[0].isspace() and
         not l.startswith(('input ', 'param ', 'local ', 'output '))]
  for stmt in program.input_stmts + program.param_stmts:
    if stmt.haoda_type.is_fixed:
      stmt = stmt.visit(lambda obj, args: obj)
      stmt.haoda_type = stmt.haoda_type.raw_type
    out.append(str(stmt))
  for kind_name, stmts in (('local', program.local_stmts),
                           ('output', program.output_stmts)):
    for stmt in stmts:
      translate = _Translator(symbols)
      for let in stmt.let:
        value, kind, frac = translate(let.expr)
        let_t = let.haoda_type
        value, kind, frac = translate.convert(value, kind, frac, let_t)
        if let_t is not None and let_t.is_fixed:
          value = '%s(%s)' % (let_t.raw_type, value)  # wrapped by widths.py
          let_t = let_t.raw_type
        elif let_t is None and kind == FIXED:
          let_t = ir.INT64
        translate.variables[let.name] = (kind, frac)
        translate.lets.append('%s%s = %s' % (
            '%s ' % let_t if let_t is not None else '', let.name, value))
      value, kind, frac = translate(stmt.expr)
      t = stmt.haoda_type
      value, kind, frac = translate.convert(value, kind, frac, t)
      if t.is_fixed:
        t = t.raw_type
      lets = translate.lets
      dram = ''
      if kind_name == 'output':
        dram = 'dram %s ' % '.'.join(map(str, stmt.dram or (0,)))
      let_text = ''.join('\n  %s' % l for l in lets)
      out.append('%s %s%s:%s %s = %s' % (kind_name, dram, t,
                                         let_text + ('\n ' if lets else ''),
                                         stmt.ref, value))
  lowered = '\n'.join(out) + '\n'
  grammar.parse(lowered)  # must be a valid program
  return lowered


def lower(stencil):
  """``Stencil`` -> ``Stencil`` without fixed-point types (the same object
  when it has none)."""
  if not has_fixed_types(stencil):
    return stencil
  from soda_b200 import sodac  # late: sodac imports the backends
  return sodac.compile_source(lower_text(str(stencil)))
