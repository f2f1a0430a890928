"""Integer widths C++ does not have (``uint6``, ``int18`` ...) on a machine that
only has 8/16/32/64-bit registers.

The reference prints such types as ``ap_uint<N>`` / ``ap_int<N>``
(haoda ``Type.c_type``): a value is *stored* in N bits - an assignment keeps
the low N bits, sign-extended for ``ap_int`` - while the arithmetic in between
is exact (``ap_int`` operators widen their result).  This pass rewrites a
program that uses such types into one that only uses <cstdint> types:

* every tensor / let / cast of type ``uintN`` (``intN``) becomes the smallest
  standard container type (``uint8``, ``uint16``, ``uint32``; ``int8`` ...;
  ``int64`` for every width of 33..63 bits, ``ir.Type.container``);
* every value stored into it is wrapped explicitly, in the DSL itself:
  ``(e) & (2^N - 1)`` for ``uintN``, ``(((e) & (2^N - 1)) ^ 2^(N-1)) - 2^(N-1)``
  for ``intN``;
* loads of *input* tensors are wrapped too (an ``ap_uint<N>`` array can only
  hold N-bit values; a caller's container array could hold more).

Exactness of the arithmetic in between is the same contract as for the
standard narrow types: intermediates must fit in 32 bits
(63 bits once a 33..63-bit value takes part: those are loaded as ``int64``),
where C++ arithmetic and the widening ``ap_int`` arithmetic agree.
Fixed-point types (``uint18_3``) are rewritten to such integers first
(fixed_point.py); custom float types stay unsupported (the reference has no C
type for them either).

The oracle does not use this pass: oracle/golden.py and oracle/emit_cpp.py
wrap natively, so that a mistake here cannot hide.
"""
from typing import Optional

from soda_b200 import grammar, ir, util


def is_narrow(t: Optional[ir.Type]) -> bool:
  return t is not None and t.is_lowerable


def has_narrow_types(program_or_stencil) -> bool:
  stmts = (list(program_or_stencil.input_stmts) +
           list(program_or_stencil.param_stmts) +
           list(program_or_stencil.local_stmts) +
           list(program_or_stencil.output_stmts))
  found = []

  def look(obj, args):
    if isinstance(obj, ir.Cast) and is_narrow(obj.haoda_type):
      found.append(obj)
    return obj

  for stmt in stmts:
    if is_narrow(stmt.haoda_type):
      return True
    for let in getattr(stmt, 'let', ()):
      if is_narrow(let.haoda_type):
        return True
      let.expr.visit(look)
    if getattr(stmt, 'expr', None) is not None:
      stmt.expr.visit(look)
  return bool(found)


def wrap_text(text: str, t: ir.Type) -> str:
  """DSL text of ``text`` wrapped to the N bits of ``t``."""
  bits = t.width_in_bits
  mask = (1 << bits) - 1
  if not t.is_signed:
    return '((%s) & %d)' % (text, mask)
  sign = 1 << (bits - 1)
  return '((((%s) & %d) ^ %d) - %d)' % (text, mask, sign, sign)


def lower_text(text: str) -> str:
  """SODA source -> equivalent source that only uses <cstdint> types."""
  program = grammar.parse(text)
  if not has_narrow_types(program):
    return text
  narrow_inputs = {
      s.name: s.haoda_type for s in program.input_stmts + program.param_stmts
      if is_narrow(s.haoda_type)
  }

  def rewrite(expr: ir.Node) -> ir.Node:
    """Casts to narrow types and loads of narrow inputs, innermost first."""

    def callback(obj, args):
      if isinstance(obj, ir.Cast) and is_narrow(obj.haoda_type):
        inner = rewrite(obj.expr)
        return grammar.parse_expr('%s(%s)' % (
            obj.haoda_type.container, wrap_text(str(inner), obj.haoda_type)))
      if isinstance(obj, ir.Ref) and obj.name in narrow_inputs:
        return grammar.parse_expr(wrap_text(str(obj), narrow_inputs[obj.name]))
      return obj

    return expr.visit(callback)

  # statements are rebuilt one by one so that the header is kept verbatim
  # (directives start in column 0; the indented lines are the let bindings
  # and the stored expression of a multi-line statement)
  out = [l for l in str(program).split('\n')
         if l.strip() and not l[0].isspace() and
         not l.startswith(('input ', 'param ', 'local ', 'output '))]
  for stmt in program.input_stmts + program.param_stmts:
    if is_narrow(stmt.haoda_type):
      stmt = stmt.visit(lambda obj, args: obj)
      stmt.haoda_type = stmt.haoda_type.container
    out.append(str(stmt))
  for kind, stmts in (('local', program.local_stmts),
                      ('output', program.output_stmts)):
    for stmt in stmts:
      t = stmt.haoda_type
      lets = []
      for let in stmt.let:
        expr = rewrite(let.expr)
        let_t = let.haoda_type
        if is_narrow(let_t):
          expr = grammar.parse_expr(wrap_text(str(expr), let_t))
          let_t = let_t.container
        lets.append('%s%s = %s' % ('%s ' % let_t if let_t is not None else '',
                                   let.name, ir.unparenthesize(expr)))
      expr = rewrite(stmt.expr)
      if is_narrow(t):
        expr = grammar.parse_expr(wrap_text(str(expr), t))
        t = t.container
      dram = ''
      if kind == 'output':
        dram = 'dram %s ' % '.'.join(map(str, stmt.dram or (0,)))
      let_text = ''.join('\n  %s' % l for l in lets)
      out.append('%s %s%s:%s %s = %s' % (kind, dram, t,
                                         let_text + ('\n ' if lets else ''),
                                         stmt.ref, ir.unparenthesize(expr)))
  lowered = '\n'.join(out) + '\n'
  grammar.parse(lowered)  # must be a valid program
  return lowered


def lower(stencil):
  """``Stencil`` -> ``Stencil`` without narrow types (the same object when it
  has none)."""
  if not has_narrow_types(stencil):
    return stencil
  from soda_b200 import sodac  # late: sodac imports the backends
  for t in (list(stencil.input_types) + list(stencil.output_types) +
            list(stencil.local_types) + list(stencil.param_types)):
    if not t.is_executable and not t.is_lowerable:
      raise util.SemanticError(
          'type %s is not supported by the CUDA backend (integers up to 64 '
          'bits, half, float and double are)' % t)
  return sodac.compile_source(lower_text(str(stencil)))
