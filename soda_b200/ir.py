"""Expression IR for SODA programs.

The reference keeps its expression grammar, node classes, type rules and C
printer in the un-vendored ``haoda`` package (reference:
src/soda/grammar.py:10-11,46,209-232 lists the classes it expects).  This module
re-states that surface from its call sites:

* node classes ``Let, Ref, Expr, LogicAnd, BinaryOr, Xor, BinaryAnd, EqCmp,
  LtCmp, AddSub, MulDiv, Unary, Operand, Cast, Call, Var`` with ``visit`` and a
  ``__str__`` that round-trips the DSL text (reference:
  src/tests/test_grammar.py:24-57);
* ``Type`` with ``c_type`` / ``width_in_bits`` / ``is_float``;
* ``propagate_type`` implementing the C++ usual arithmetic conversions, because
  the reference's results are whatever g++ computes for the printed expression
  (SURVEY.md appendix A.3);
* a C/CUDA expression printer shared by nothing else: the oracle has its own
  evaluator so that a printer bug cannot hide.
"""
import copy
import re
from typing import Callable, Dict, Iterable, List, Optional, Tuple

from soda_b200 import util

# ---------------------------------------------------------------------------
# types
# ---------------------------------------------------------------------------

_TYPE_RE = re.compile(r'^(u?int|float)(\d+)(?:_(\d+))?$')
_STD_INT_WIDTHS = (8, 16, 32, 64)


class Type:
  """A SODA element type such as ``uint16``, ``int32``, ``float``, ``double``.

  Arbitrary widths (``uint6``, ``int27``, ``float18_3``) parse and print.  The
  widths C++ has (<cstdint> integers, float, double) and IEEE ``half`` execute
  as such; other integers of up to 63 bits are lowered to containers
  (``is_lowerable``), fixed-point types to scaled integers
  (soda_b200/optimization/fixed_point.py); ``c_type`` raises for what remains:
  custom floats, which the reference cannot print as C either.
  """
  __slots__ = ('_name',)

  def __init__(self, name):
    if isinstance(name, Type):
      name = name._name
    self._name = str(name)

  def __str__(self) -> str:
    return self._name

  def __repr__(self) -> str:
    return 'Type(%r)' % self._name

  def __eq__(self, other) -> bool:
    if isinstance(other, Type):
      return self._name == other._name
    if isinstance(other, str):
      return self._name == other
    return NotImplemented

  def __ne__(self, other) -> bool:
    result = self.__eq__(other)
    return result if result is NotImplemented else not result

  def __hash__(self) -> int:
    return hash(self._name)

  @property
  def is_float(self) -> bool:
    return self._name in ('float', 'double', 'half') or \
        self._name.startswith('float')

  @property
  def is_signed(self) -> bool:
    return self.is_float or self._name.startswith('int')

  @property
  def is_fixed(self) -> bool:
    m = _TYPE_RE.match(self._name)
    return bool(m and m.group(1) != 'float' and m.group(3) is not None)

  @property
  def frac_bits(self) -> int:
    """Fractional bits of a fixed-point type ``(u)intN_M`` (N bits in all, M of
    them integer bits: the reference prints ``ap_(u)fixed<N, M>``); 0 for
    every other type."""
    m = _TYPE_RE.match(self._name)
    if m and m.group(1) != 'float' and m.group(3) is not None:
      return int(m.group(2)) - int(m.group(3))
    return 0

  @property
  def raw_type(self) -> 'Type':
    """The integer type of the same width that holds the scaled value of a
    fixed-point type (``uint18_3`` -> ``uint18``)."""
    if not self.is_fixed:
      return self
    m = _TYPE_RE.match(self._name)
    return Type(m.group(1) + m.group(2))

  @staticmethod
  def exact_fixed(frac_bits: int) -> 'Type':
    """Type of an exact fixed-point intermediate with ``frac_bits`` fractional
    bits (``ap_fixed`` operators widen; here: 64 bits)."""
    return Type('int64_%d' % (64 - frac_bits))

  @property
  def width_in_bits(self) -> int:
    if self._name == 'float':
      return 32
    if self._name == 'double':
      return 64
    if self._name == 'half':
      return 16
    if self._name == 'bool':
      return 1
    m = _TYPE_RE.match(self._name)
    if m is None:
      raise util.InternalError('unknown type: %s' % self._name)
    return int(m.group(2))

  @property
  def is_executable(self) -> bool:
    """Whether C++ has this exact type (so oracle and GPU agree bit for bit)."""
    if self._name in ('float', 'double', 'half', 'bool'):
      return True
    m = _TYPE_RE.match(self._name)
    return bool(m and m.group(1) != 'float' and m.group(3) is None and
                int(m.group(2)) in _STD_INT_WIDTHS)

  @property
  def is_lowerable(self) -> bool:
    """An integer of 1..63 bits that C++ does not have: stored in the next
    <cstdint> container and wrapped to its width on every store
    (soda_b200/optimization/widths.py)."""
    m = _TYPE_RE.match(self._name)
    return bool(m and m.group(1) != 'float' and m.group(3) is None and
                int(m.group(2)) not in _STD_INT_WIDTHS and
                1 <= int(m.group(2)) <= 63)

  @property
  def container(self) -> 'Type':
    """The <cstdint> type that stores a value of this type.  33..63-bit values,
    signed or not, live in an ``int64``: it holds every such value, and the
    arithmetic on what is loaded from it is then signed 64-bit - the widening
    arithmetic of ``ap_int`` as long as intermediates fit 63 bits (an
    ``ap_uint<40>`` occupies 8 bytes with the value zero-extended, so the bytes
    are those of the reference's array)."""
    if self.is_fixed:
      return self.raw_type.container
    if not self.is_lowerable:
      return self
    bits = self.width_in_bits
    if bits > 32:
      return Type('int64')
    width = 8 if bits <= 8 else 16 if bits <= 16 else 32
    return Type('%s%d' % ('int' if self.is_signed else 'uint', width))

  @property
  def c_type(self) -> str:
    if self._name in ('float', 'double', 'bool'):
      return self._name
    if self._name == 'half':
      return 'soda::half_t'  # soda_b200/csrc/soda_half.cuh
    if not self.is_executable:
      raise util.SemanticError(
          'type %s has no C++ equivalent; the CUDA backend and the oracle '
          'support uint8/16/32/64, int8/16/32/64, half, float and double' %
          self._name)
    return self._name + '_t'

  @property
  def numpy_name(self) -> str:
    if self._name == 'float':
      return 'float32'
    if self._name == 'double':
      return 'float64'
    if self._name == 'half':
      return 'float16'
    if self._name == 'bool':
      return 'bool'
    if not self.is_executable:
      raise util.SemanticError('type %s has no numpy equivalent' % self._name)
    return self._name


def as_type(value) -> Optional[Type]:
  if value is None or isinstance(value, Type):
    return value
  return Type(value)


INT32 = Type('int32')
UINT32 = Type('uint32')
INT64 = Type('int64')
UINT64 = Type('uint64')
FLOAT = Type('float')
DOUBLE = Type('double')
HALF = Type('half')


def _promote(t: Type) -> Type:
  """C++ integral promotion."""
  if t.is_float:
    return t
  if t == 'bool' or t.width_in_bits < 32:
    return INT32
  if t.is_lowerable:
    return INT64  # 33..63 bits: see Type.container
  return t


def common_type(a: Type, b: Type) -> Type:
  """C++ usual arithmetic conversions for two executable types."""
  # half ranks below float and above every integer type, like _Float16
  for t in (DOUBLE, FLOAT, HALF):
    if a == t or b == t:
      return t
  a, b = _promote(a), _promote(b)
  if a == b:
    return a
  wa, wb = a.width_in_bits, b.width_in_bits
  if a.is_signed == b.is_signed:
    return a if wa >= wb else b
  signed, unsigned = (a, b) if a.is_signed else (b, a)
  if unsigned.width_in_bits >= signed.width_in_bits:
    return unsigned
  return signed


# ---------------------------------------------------------------------------
# nodes
# ---------------------------------------------------------------------------


class Node:
  """Base class: attributes are declared in SCALAR_ATTRS / LINEAR_ATTRS.

  Scalar attributes hold one value (possibly another Node), linear attributes
  hold tuples.  Equality and hashing are structural over those attributes.
  """
  SCALAR_ATTRS: Tuple[str, ...] = ()
  LINEAR_ATTRS: Tuple[str, ...] = ()

  def __init__(self, **kwargs):
    for attr in self.SCALAR_ATTRS:
      setattr(self, attr, kwargs.pop(attr, None))
    for attr in self.LINEAR_ATTRS:
      value = kwargs.pop(attr, ())
      setattr(self, attr, tuple(value) if value is not None else ())
    self._tx_position = kwargs.pop('_tx_position', 0)
    if 'haoda_type' in self.SCALAR_ATTRS:
      self.haoda_type = as_type(self.haoda_type)
    else:
      # derived / propagated type of an expression node
      self.haoda_type = as_type(kwargs.pop('haoda_type', None))

  @property
  def ATTRS(self) -> Tuple[str, ...]:
    return self.SCALAR_ATTRS + self.LINEAR_ATTRS

  def _key(self):
    return (type(self).__name__,
            tuple(getattr(self, a) for a in self.SCALAR_ATTRS),
            tuple(tuple(getattr(self, a)) for a in self.LINEAR_ATTRS))

  def __hash__(self) -> int:
    return hash(self._key())

  def __eq__(self, other) -> bool:
    if other is None or type(other) is not type(self):
      return False
    return self._key() == other._key()

  def __ne__(self, other) -> bool:
    return not self.__eq__(other)

  @property
  def c_type(self) -> str:
    return self.haoda_type.c_type

  @property
  def width_in_bits(self) -> int:
    return self.haoda_type.width_in_bits

  def visit(self,
            callback: Optional[Callable] = None,
            args=None,
            pre_recursion: Optional[Callable] = None,
            post_recursion: Optional[Callable] = None):
    """Functional tree walk that returns a (possibly) rewritten copy.

    ``callback(node_copy, args)`` may mutate and/or return a replacement.  If it
    returns a different object, that object replaces the subtree without further
    recursion; otherwise children that the callback left untouched are visited
    recursively.  The receiver itself is never modified.
    """

    def call(func, obj):
      if func is None:
        return obj
      result = func(obj, args)
      return obj if result is None else result

    mine = copy.copy(self)
    obj = call(callback, mine)
    if obj is not mine:
      return obj
    call(pre_recursion, obj)
    for attr in self.SCALAR_ATTRS:
      child = getattr(obj, attr, None)
      if child is getattr(self, attr, None) and isinstance(child, Node):
        setattr(obj, attr,
                child.visit(callback, args, pre_recursion, post_recursion))
    for attr in self.LINEAR_ATTRS:
      children = getattr(obj, attr, None)
      if children is getattr(self, attr, None) and children:
        setattr(
            obj, attr,
            tuple(
                c.visit(callback, args, pre_recursion, post_recursion
                       ) if isinstance(c, Node) else c for c in children))
    return call(post_recursion, obj)


class Let(Node):
  SCALAR_ATTRS = 'haoda_type', 'name', 'expr'

  def __str__(self) -> str:
    result = '{} = {}'.format(self.name, unparenthesize(self.expr))
    if self.haoda_type is not None:
      result = '{} {}'.format(self.haoda_type, result)
    return result


class Ref(Node):
  """``name(i, j, ...)`` with an optional ``~latency`` annotation."""
  SCALAR_ATTRS = 'name', 'lat'
  LINEAR_ATTRS = ('idx',)

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    self.idx = tuple(int(x) for x in self.idx)
    self.parent = kwargs.get('parent')

  def __str__(self) -> str:
    result = '{}({})'.format(self.name, ', '.join(map(str, self.idx)))
    if self.lat is not None:
      result += ' ~{}'.format(self.lat)
    return result


class BinaryOp(Node):
  """N-ary chain ``operand[0] (operator[i] operand[i+1])*`` at one precedence
  level, evaluated left to right."""
  LINEAR_ATTRS = 'operand', 'operator'

  def __str__(self) -> str:
    result = str(self.operand[0])
    for operator, operand in zip(self.operator, self.operand[1:]):
      result += ' {} {}'.format(operator, operand)
    return result

  @property
  def singleton(self) -> bool:
    return len(self.operand) == 1


class Expr(BinaryOp):  # ||
  pass


class LogicAnd(BinaryOp):  # &&
  pass


class BinaryOr(BinaryOp):  # |
  pass


class Xor(BinaryOp):  # ^
  pass


class BinaryAnd(BinaryOp):  # &
  pass


class EqCmp(BinaryOp):  # == !=
  pass


class LtCmp(BinaryOp):  # <= >= < >
  pass


class AddSub(BinaryOp):  # + -
  pass


class MulDiv(BinaryOp):  # * / %
  pass


class Unary(Node):
  SCALAR_ATTRS = ('operand',)
  LINEAR_ATTRS = ('operator',)

  def __str__(self) -> str:
    return ''.join(self.operator) + str(self.operand)


class Operand(Node):
  """A leaf or a parenthesised sub-expression (``expr``)."""
  SCALAR_ATTRS = 'cast', 'call', 'ref', 'num', 'var', 'expr'

  def __str__(self) -> str:
    for attr in ('cast', 'call', 'ref', 'num', 'var'):
      value = getattr(self, attr)
      if value is not None:
        return str(value)
    return '(%s)' % str(self.expr)

  @property
  def inner(self):
    for attr in self.SCALAR_ATTRS:
      value = getattr(self, attr)
      if value is not None:
        return value
    raise util.InternalError('empty operand')


class Cast(Node):
  SCALAR_ATTRS = 'haoda_type', 'expr'

  def __str__(self) -> str:
    return '{}({})'.format(self.haoda_type, unparenthesize(self.expr))


class Call(Node):
  SCALAR_ATTRS = ('name',)
  LINEAR_ATTRS = ('arg',)

  def __str__(self) -> str:
    return '{}({})'.format(self.name, ', '.join(map(str, self.arg)))


class Var(Node):
  """A ``let`` variable or a ``param`` element: ``name`` or ``name[i][j]``."""
  SCALAR_ATTRS = ('name',)
  LINEAR_ATTRS = ('idx',)

  def __str__(self) -> str:
    return self.name + ''.join('[%d]' % x for x in self.idx)


class Num(Node):
  """A numeric literal, kept as source text so that it prints back verbatim and
  is typed like the C++ literal it is (``3`` int, ``0.2f`` float, ``2.0``
  double, ``7u`` unsigned)."""
  SCALAR_ATTRS = ('text',)

  def __str__(self) -> str:
    return self.text

  @property
  def literal_type(self) -> Type:
    text = self.text.lower()
    is_hex = text.startswith('0x')
    if not is_hex and (any(c in text for c in '.e') or text.endswith('f')):
      return FLOAT if text.endswith('f') else DOUBLE
    suffix = text.lstrip('0123456789abcdefx')
    value = self.value
    if 'u' in suffix:
      if 'l' in suffix or value >= 2**32:
        return UINT64
      return UINT32
    if 'l' in suffix or value >= 2**31:
      # decimal literals never become unsigned int; hex/octal ones may
      if (is_hex or (text.startswith('0') and len(text) > 1)) and \
          'l' not in suffix and value < 2**32:
        return UINT32
      return INT64
    return INT32

  @property
  def value(self):
    text = self.text.lower()
    if text.startswith('0x'):
      return int(text.rstrip('ul'), 16)
    if any(c in text for c in '.e') or text.endswith('f'):
      return float(text.rstrip('f'))
    text = text.rstrip('ul')
    if text.startswith('0b'):
      return int(text, 2)
    if text.startswith('0') and len(text) > 1:
      return int(text, 8)
    return int(text)

  @property
  def c_literal(self) -> str:
    text = self.text
    if text.lower().startswith('0b'):
      return str(self.value)
    return text


def unparenthesize(node):
  """Strip redundant outer parentheses: ``((a + b))`` -> ``a + b``."""
  while isinstance(node, Operand) and node.expr is not None:
    node = node.expr
  return node


def unwrap(node):
  """Strip Operand wrappers and single-operand chains around a node."""
  while True:
    if isinstance(node, Operand):
      node = node.inner
    elif isinstance(node, BinaryOp) and node.singleton:
      node = node.operand[0]
    elif isinstance(node, Unary) and not node.operator:
      node = node.operand
    else:
      return node


def make_var(name: str, haoda_type=None) -> Var:
  var = Var(name=name, idx=())
  var.haoda_type = haoda_type
  return var


def get_vars(node: Node) -> Tuple[Var, ...]:
  found: List[Var] = []

  def visitor(obj, args):
    if isinstance(obj, Var):
      found.append(obj)

  node.visit(visitor)
  return tuple(found)


# ---------------------------------------------------------------------------
# type propagation (C++ usual arithmetic conversions)
# ---------------------------------------------------------------------------

# Calls that g++ resolves to the C library's double-precision function when the
# generated code says e.g. ``sqrt(x)`` with only <cmath> included (verified
# with g++ 13.3: decltype(sqrt(1.0f)) is double).
DOUBLE_MATH_CALLS = ('sqrt', 'exp', 'log', 'fabs', 'floor', 'ceil', 'pow',
                     'sin', 'cos', 'tanh')
# ... unless the float overloads are visible in the global namespace: the
# reference's generated files also include Xilinx's <ap_int.h>
# (src/soda/codegen/frt/host.py:33, xilinx/hls_kernel.py:228-233), which is not
# on disk here and may pull ``std::sqrt(float)`` and friends into scope.  The
# ``float`` math mode (sodac --math-precision float) models that reading: these
# calls - the ones that are correctly rounded in binary32 on both the CPU and
# the GPU - keep a float argument in float.  The default stays double, the
# behaviour that can be checked here (plain <cmath>, g++ 13.3).
FLOAT_MATH_CALLS = ('sqrt', 'fabs', 'floor', 'ceil')
SELECT_CALLS = ('min', 'max')


def result_type(node: Node) -> Type:
  """Type of an already-propagated node."""
  t = node.haoda_type
  if t is None:
    raise util.InternalError('type of `%s` was not propagated' % node)
  return t


def propagate_type(node: Node, symbol_table: Dict[str, Type],
                   float_math: bool = False) -> Node:
  """Returns a copy of ``node`` with ``haoda_type`` set on every sub-node.

  ``symbol_table`` maps tensor names and let-variable names to their types.
  ``float_math``: ``sqrt`` / ``fabs`` / ``floor`` / ``ceil`` of a ``float``
  stay ``float`` (see ``FLOAT_MATH_CALLS``) instead of going through double.
  """

  def post(obj, args):
    if isinstance(obj, Ref):
      if obj.name not in symbol_table:
        raise util.SemanticError('undefined tensor `%s`' % obj.name)
      obj.haoda_type = symbol_table[obj.name]
    elif isinstance(obj, Var):
      if obj.name not in symbol_table:
        raise util.SemanticError('undefined variable `%s`' % obj.name)
      obj.haoda_type = symbol_table[obj.name]
    elif isinstance(obj, Num):
      obj.haoda_type = obj.literal_type
    elif isinstance(obj, Operand):
      obj.haoda_type = result_type(obj.inner)
    elif isinstance(obj, (Cast, Let)):
      if isinstance(obj, Let) and obj.haoda_type is None:
        obj.haoda_type = result_type(obj.expr)
    elif isinstance(obj, Unary):
      t = result_type(obj.operand)
      for op in reversed(obj.operator):
        t = Type('bool') if op == '!' else _promote(t)
      obj.haoda_type = t
    elif isinstance(obj, (Expr, LogicAnd, EqCmp, LtCmp)):
      obj.haoda_type = (result_type(obj.operand[0])
                        if obj.singleton else Type('bool'))
    elif isinstance(obj, BinaryOp):
      t = result_type(obj.operand[0])
      for opd in obj.operand[1:]:
        t = common_type(t, result_type(opd))
      obj.haoda_type = t
    elif isinstance(obj, Call):
      if obj.name in DOUBLE_MATH_CALLS:
        if float_math and obj.name in FLOAT_MATH_CALLS and all(
            result_type(arg) == FLOAT for arg in obj.arg):
          obj.haoda_type = FLOAT
        else:
          obj.haoda_type = DOUBLE
      elif obj.name in SELECT_CALLS:
        t = result_type(obj.arg[0])
        for arg in obj.arg[1:]:
          t = common_type(t, result_type(arg))
        obj.haoda_type = _promote(t) if len(obj.arg) > 1 else t
      elif obj.name == 'abs':
        obj.haoda_type = _promote(result_type(obj.arg[0]))
      else:
        raise util.SemanticError('unsupported function `%s`' % obj.name)
    return obj

  return node.visit(post_recursion=post)


# ---------------------------------------------------------------------------
# C++ / CUDA expression printer
# ---------------------------------------------------------------------------


class CPrinter:
  """Prints a type-propagated expression as a fully parenthesised C++ expression.

  ``ref_printer(ref)`` renders a tensor access, ``var_printer(var)`` a variable.
  Every n-ary chain is printed left-to-right inside one pair of parentheses, so
  C++ evaluates it in IR order; no re-association, no constant folding.
  """

  def __init__(self, ref_printer: Callable[[Ref], str],
               var_printer: Optional[Callable[[Var], str]] = None,
               min_name: str = 'soda_min', max_name: str = 'soda_max'):
    self.ref_printer = ref_printer
    self.var_printer = var_printer or (lambda var: str(var))
    self.min_name = min_name
    self.max_name = max_name

  def __call__(self, node: Node) -> str:
    if isinstance(node, Operand):
      return self(node.inner)
    if isinstance(node, Ref):
      return self.ref_printer(node)
    if isinstance(node, Var):
      return self.var_printer(node)
    if isinstance(node, Num):
      return node.c_literal
    if isinstance(node, Cast):
      return '{}({})'.format(node.haoda_type.c_type, self(node.expr))
    if isinstance(node, Unary):
      # a space keeps `- -x` from lexing as `--x`
      return '(' + ''.join(op + ' ' for op in node.operator) + \
          self(node.operand) + ')'
    if isinstance(node, BinaryOp):
      if node.singleton:
        return self(node.operand[0])
      result = self(node.operand[0])
      for operator, operand in zip(node.operator, node.operand[1:]):
        result += ' {} {}'.format(operator, self(operand))
      return '(' + result + ')'
    if isinstance(node, Call):
      args = [self(arg) for arg in node.arg]
      if node.name in SELECT_CALLS:
        name = self.min_name if node.name == 'min' else self.max_name
        ctype = result_type(node).c_type
        result = '{}({})'.format(ctype, args[0])
        for arg in args[1:]:
          result = '{}<{}>({}, {}({}))'.format(name, ctype, result, ctype, arg)
        return result
      if node.name in DOUBLE_MATH_CALLS:
        if node.haoda_type == FLOAT:  # float math mode
          return '{}f(float({}))'.format(node.name, args[0])
        return '{}(double({}){})'.format(
            node.name, args[0],
            ''.join(', double(%s)' % a for a in args[1:]))
      if node.name == 'abs':
        ctype = result_type(node).c_type
        return 'soda_abs<{}>({}({}))'.format(ctype, ctype, args[0])
      raise util.SemanticError('unsupported function `%s`' % node.name)
    raise util.InternalError('cannot print %r' % node)


# ---------------------------------------------------------------------------
# reductions (used by rebalance / computation reuse)
# ---------------------------------------------------------------------------


def to_reduction(expr: Node) -> Optional[Tuple]:
  """Flattens ``a + (b + c)``-style trees into ``('+', (a, b, c))``.

  Returns None if ``expr`` is not a pure ``+`` or pure ``*`` reduction with at
  least two operands.  (reference call sites:
  src/soda/optimization/computation_reuse.py:730,1792.)
  """
  expr = unwrap(expr)
  if isinstance(expr, AddSub):
    op = '+'
  elif isinstance(expr, MulDiv):
    op = '*'
  else:
    return None
  if set(expr.operator) != {op}:
    return None

  def flatten(node) -> Iterable[Node]:
    node = unwrap(node)
    if type(node) is type(expr) and set(node.operator) == {op}:
      for operand in node.operand:
        yield from flatten(operand)
    else:
      yield node

  operands = tuple(flatten(expr))
  if len(operands) < 2:
    return None
  return op, operands


def from_reduction(op: str, operands: Iterable[Node]) -> Node:
  operands = tuple(operands)
  cls = AddSub if op == '+' else MulDiv
  if len(operands) == 1:
    return operands[0]
  wrapped = tuple(
      Operand(expr=o) if isinstance(o, BinaryOp) and not o.singleton else o
      for o in operands)
  return cls(operand=wrapped, operator=(op,) * (len(wrapped) - 1))
