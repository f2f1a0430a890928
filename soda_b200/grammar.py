"""SODA DSL front end: lexer, recursive-descent parser and statement nodes.

The reference builds its parser with textX from a grammar string whose
expression half lives in the un-vendored ``haoda`` package (reference:
src/soda/grammar.py:15-46).  Neither textX nor haoda exist in this image, so the
grammar is implemented by hand here.  What is kept identical is the language
accepted (header directives in any order, ``input``/``param``/``local``/
``output`` statements, ``let`` bindings, ``~latency`` annotations, comments) and
the ``__str__`` round trip pinned by src/tests/test_grammar.py:24-137.

Statement classes keep the reference's attribute names (``haoda_type``,
``name``, ``tile_size``, ``dram``, ``ref``, ``expr``, ``let``) so that
``core.Stencil(**program.__dict__)`` style construction
(src/tests/test_core.py:58-60) works unchanged.
"""
import re
from typing import Dict, List, Optional, Tuple

from soda_b200 import ir, util

FUNC_NAMES = frozenset(ir.DOUBLE_MATH_CALLS + ir.SELECT_CALLS + ('abs',))

_TYPE_NAME_RE = re.compile(
    r'^(?:u?int\d+(?:_\d+)?|float\d+(?:_\d+)?|float|double|half)$')

_TOKEN_RE = re.compile(
    r'''
    (?P<ws>\s+)
  | (?P<comment>\#[^\n]*)
  | (?P<num>
        0[xX][0-9a-fA-F]+[uUlL]*
      | 0[bB][01]+[uUlL]*
      | (?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fF]?[uUlL]*
    )
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>\|\||&&|==|!=|<=|>=|[-+*/%<>!~^&|=(){}\[\],.:])
    ''', re.VERBOSE)


class SyntaxError_(util.SemanticError):
  """Raised on malformed SODA source (the reference raises TextXSyntaxError)."""


class _Token:
  __slots__ = ('kind', 'text', 'pos')

  def __init__(self, kind: str, text: str, pos: int):
    self.kind, self.text, self.pos = kind, text, pos

  def __repr__(self):
    return '%s:%r@%d' % (self.kind, self.text, self.pos)


def _tokenize(text: str) -> List[_Token]:
  tokens = []
  pos = 0
  while pos < len(text):
    m = _TOKEN_RE.match(text, pos)
    if m is None:
      line = text.count('\n', 0, pos) + 1
      raise SyntaxError_('unexpected character %r at line %d' %
                         (text[pos], line))
    kind = m.lastgroup
    if kind not in ('ws', 'comment'):
      tokens.append(_Token(kind, m.group(kind), pos))
    pos = m.end()
  tokens.append(_Token('eof', '', len(text)))
  return tokens


# ---------------------------------------------------------------------------
# statement nodes
# ---------------------------------------------------------------------------


class InputStmt(ir.Node):
  """``input [dram B(.B)*] Type: name[(T0, T1, ..., *)]``.

  ``tile_size`` always ends with a 0 for the unbounded last dimension
  (reference: src/soda/grammar.py:48-71).
  """
  SCALAR_ATTRS = 'haoda_type', 'name'
  LINEAR_ATTRS = 'tile_size', 'dram'

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    if not self.dram:
      self.dram = (0,)
    self.tile_size = tuple(self.tile_size) + (0,)

  def __str__(self) -> str:
    result = 'input dram {} {}: {}'.format('.'.join(map(str, self.dram)),
                                           self.haoda_type, self.name)
    if self.tile_size[:-1]:
      result += '({}, *)'.format(', '.join(map(str, self.tile_size[:-1])))
    return result


class LocalStmtOrOutputStmt(ir.Node):
  SCALAR_ATTRS = 'haoda_type', 'ref', 'expr'
  LINEAR_ATTRS: Tuple[str, ...] = ('let',)

  def __init__(self, **kwargs):
    self.stencil = kwargs.pop('stencil', None)
    super().__init__(**kwargs)
    var_types = {let.name: let.haoda_type for let in self.let}

    def set_var_type(obj, args):
      if isinstance(obj, ir.Var) and obj.name in var_types:
        obj.haoda_type = var_types[obj.name]
      return obj

    self.let = tuple(let.visit(set_var_type) for let in self.let)
    if self.expr is not None:
      self.expr = self.expr.visit(set_var_type)

  @property
  def name(self) -> str:
    return self.ref.name

  def __str__(self) -> str:
    if self.let:
      let = '\n  {}\n '.format('\n  '.join(map(str, self.let)))
    else:
      let = ''
    return '{}:{} {} = {}'.format(self.haoda_type, let, self.ref,
                                  ir.unparenthesize(self.expr))

  @property
  def symbol_table(self) -> Dict[str, ir.Type]:
    """Tensor types plus the types of this statement's own let variables
    (reference: src/soda/grammar.py:108-121).  Lets are resolved in dependency
    order so a let may use an earlier (or later-defined) let."""
    table = dict(self.stencil.symbol_table) if self.stencil is not None else {}
    pending = {let.name: let for let in self.let}
    progress = True
    while pending and progress:
      progress = False
      for name, let in list(pending.items()):
        deps = {v.name for v in ir.get_vars(let.expr)} & set(pending)
        if deps - {name}:
          continue
        if let.haoda_type is not None:
          table[name] = let.haoda_type
        else:
          table[name] = ir.propagate_type(let.expr, table,
                                          self._float_math).haoda_type
        del pending[name]
        progress = True
    if pending:
      raise util.SemanticError('circular let bindings: %s' %
                               ', '.join(sorted(pending)))
    return table

  def propagate_type(self, dummy=None) -> None:
    """Type every node; wrap the expression in a Cast to the declared tensor
    type if it differs (reference: src/soda/grammar.py:123-136)."""
    table = self.symbol_table
    self.expr = ir.propagate_type(self.expr, table, self._float_math)
    if self.expr.haoda_type != self.haoda_type:
      self.expr = ir.Cast(expr=self.expr, haoda_type=self.haoda_type)
    self.let = tuple(ir.propagate_type(let, table, self._float_math)
                     for let in self.let)

  @property
  def _float_math(self) -> bool:
    return bool(getattr(getattr(self, 'stencil', None), 'float_math', False))


class LocalStmt(LocalStmtOrOutputStmt):

  def __str__(self) -> str:
    return 'local ' + super().__str__()


class OutputStmt(LocalStmtOrOutputStmt):
  LINEAR_ATTRS = LocalStmtOrOutputStmt.LINEAR_ATTRS + ('dram',)

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    if not self.dram:
      self.dram = (0,)

  def __str__(self) -> str:
    return 'output dram {} {}'.format('.'.join(map(str, self.dram)),
                                      super().__str__())


class Partitioning(ir.Node):
  SCALAR_ATTRS = 'strategy', 'factor', 'dim'


class ParamAttr(ir.Node):
  SCALAR_ATTRS = 'dup', 'partitioning'

  def __str__(self) -> str:
    if self.dup is not None:
      return 'dup {}'.format(self.dup)
    result = 'partition {}'.format(self.partitioning.strategy)
    if self.partitioning.strategy == 'cyclic':
      result += ' factor={}'.format(self.partitioning.factor)
    if self.partitioning.dim is not None:
      result += ' dim={}'.format(self.partitioning.dim)
    return result


class ParamStmt(ir.Node):
  SCALAR_ATTRS = 'haoda_type', 'name'
  LINEAR_ATTRS = 'attr', 'size', 'dram'

  def __str__(self) -> str:
    return 'param {}{}: {}{}'.format(self.haoda_type,
                                     ''.join(map(', {}'.format, self.attr)),
                                     self.name,
                                     ''.join(map('[{}]'.format, self.size)))


class SodaProgram(ir.Node):
  SCALAR_ATTRS = ('border', 'burst_width', 'cluster', 'iterate', 'app_name',
                  'unroll_factor', 'input_stmts', 'param_stmts', 'local_stmts',
                  'output_stmts')

  def __init__(self, **kwargs):
    super().__init__(**kwargs)
    del self.haoda_type
    for attr in ('input_stmts', 'param_stmts', 'local_stmts', 'output_stmts'):
      setattr(self, attr, list(getattr(self, attr) or ()))
    # the single tiled input fixes tile sizes and dimensionality
    # (reference: src/soda/grammar.py:177-194)
    node = None
    for node in self.input_stmts:
      if hasattr(self, 'tile_size'):
        if node.tile_size[:-1] and self.tile_size != node.tile_size:
          raise util.SemanticError(
              'tile size %s doesn\'t match previous one %s' %
              (node.tile_size, self.tile_size))
      elif node.tile_size[:-1]:
        self.tile_size = node.tile_size
        self.dim = len(self.tile_size)
    if not hasattr(self, 'tile_size') and node is not None:
      self.tile_size = node.tile_size  # 1-D program
      self.dim = len(self.tile_size)

  def __str__(self) -> str:
    return '\n'.join(
        filter(None, (
            'border: {}'.format(self.border) if self.border else '',
            'burst width: {}'.format(self.burst_width),
            'cluster: {}'.format(self.cluster) if self.cluster else '',
            'iterate: {}'.format(self.iterate),
            'kernel: {}'.format(self.app_name),
            'unroll factor: {}'.format(self.unroll_factor),
            '\n'.join(map(str, self.input_stmts)),
            '\n'.join(map(str, self.param_stmts)),
            '\n'.join(map(str, self.local_stmts)),
            '\n'.join(map(str, self.output_stmts)),
        )))


# ---------------------------------------------------------------------------
# parser
# ---------------------------------------------------------------------------


class _Parser:

  def __init__(self, text: str):
    self.text = text
    self.tokens = _tokenize(text)
    self.i = 0

  # -- token helpers --------------------------------------------------------
  @property
  def tok(self) -> _Token:
    return self.tokens[self.i]

  def peek(self, offset: int = 1) -> _Token:
    return self.tokens[min(self.i + offset, len(self.tokens) - 1)]

  def error(self, what: str):
    tok = self.tok
    line = self.text.count('\n', 0, tok.pos) + 1
    col = tok.pos - (self.text.rfind('\n', 0, tok.pos) + 1) + 1
    raise SyntaxError_('line %d col %d: expected %s, got %r' %
                       (line, col, what, tok.text or 'end of input'))

  def at(self, text: str) -> bool:
    return self.tok.text == text and self.tok.kind in ('id', 'op')

  def accept(self, text: str) -> bool:
    if self.at(text):
      self.i += 1
      return True
    return False

  def expect(self, text: str) -> None:
    if not self.accept(text):
      self.error(repr(text))

  def ident(self) -> str:
    if self.tok.kind != 'id':
      self.error('an identifier')
    self.i += 1
    return self.tokens[self.i - 1].text

  def integer(self, signed: bool = False) -> int:
    sign = 1
    if signed:
      while self.tok.text in ('+', '-') and self.tok.kind == 'op':
        if self.tok.text == '-':
          sign = -sign
        self.i += 1
    if self.tok.kind != 'num' or not re.match(r'^(0[xX][0-9a-fA-F]+|\d+)$',
                                              self.tok.text):
      self.error('an integer')
    self.i += 1
    return sign * int(self.tokens[self.i - 1].text, 0)

  def is_type(self, tok: Optional[_Token] = None) -> bool:
    tok = tok or self.tok
    return tok.kind == 'id' and bool(_TYPE_NAME_RE.match(tok.text))

  def type_(self) -> ir.Type:
    if not self.is_type():
      self.error('a type')
    self.i += 1
    return ir.Type(self.tokens[self.i - 1].text)

  # -- program --------------------------------------------------------------
  def program(self) -> SodaProgram:
    header: Dict[str, object] = {}
    stmts: Dict[str, list] = {
        'input_stmts': [],
        'param_stmts': [],
        'local_stmts': [],
        'output_stmts': [],
    }

    def set_header(key, value):
      if key in header:
        raise SyntaxError_('duplicate directive `%s`' % key.replace('_', ' '))
      header[key] = value

    while self.tok.kind != 'eof':
      pos = self.tok.pos
      if self.accept('kernel'):
        self.expect(':')
        set_header('app_name', self.ident())
      elif self.accept('burst'):
        self.expect('width')
        self.expect(':')
        set_header('burst_width', self.integer())
      elif self.accept('unroll'):
        self.expect('factor')
        self.expect(':')
        set_header('unroll_factor', self.integer())
      elif self.accept('iterate'):
        self.expect(':')
        set_header('iterate', self.integer())
      elif self.accept('border'):
        self.expect(':')
        value = self.ident()
        if value not in ('ignore', 'preserve'):
          raise SyntaxError_('border must be ignore or preserve, got %s' %
                             value)
        set_header('border', value)
      elif self.accept('cluster'):
        self.expect(':')
        value = self.ident()
        if value not in ('none', 'fine', 'coarse', 'full'):
          raise SyntaxError_('unknown cluster strategy %s' % value)
        set_header('cluster', value)
      elif self.accept('input'):
        stmt = self.input_stmt()
        stmt._tx_position = pos
        stmts['input_stmts'].append(stmt)
      elif self.accept('param'):
        stmt = self.param_stmt()
        stmt._tx_position = pos
        stmts['param_stmts'].append(stmt)
      elif self.accept('local'):
        stmt = self.compute_stmt(LocalStmt)
        stmt._tx_position = pos
        stmts['local_stmts'].append(stmt)
      elif self.accept('output'):
        stmt = self.compute_stmt(OutputStmt)
        stmt._tx_position = pos
        stmts['output_stmts'].append(stmt)
      else:
        self.error('a directive or a statement')

    for key in ('burst_width', 'iterate', 'app_name', 'unroll_factor'):
      if key not in header:
        raise SyntaxError_('missing directive `%s`' %
                           {'app_name': 'kernel'}.get(key,
                                                      key.replace('_', ' ')))
    if not stmts['input_stmts']:
      raise SyntaxError_('a SODA program needs at least one input')
    if not stmts['output_stmts']:
      raise SyntaxError_('a SODA program needs at least one output')
    header.setdefault('border', None)
    header.setdefault('cluster', None)
    return SodaProgram(**header, **stmts)

  def dram(self) -> Tuple[int, ...]:
    if not self.accept('dram'):
      return ()
    # `0.1.2` lexes as the numbers `0.1` and `.2`: glue the pieces back
    text = ''
    while self.tok.kind == 'num' or self.at('.'):
      text += self.tok.text
      self.i += 1
    if not re.match(r'^\d+(\.\d+)*$', text):
      self.error('a dram bank list such as 0 or 0.1')
    return tuple(int(x) for x in text.split('.'))

  def input_stmt(self) -> InputStmt:
    dram = self.dram()
    haoda_type = self.type_()
    self.expect(':')
    name = self.ident()
    tile_size: List[int] = []
    if self.accept('('):
      while not self.accept('*'):
        tile_size.append(self.integer())
        self.expect(',')
      self.expect(')')
    return InputStmt(haoda_type=haoda_type,
                     name=name,
                     tile_size=tile_size,
                     dram=dram)

  def param_stmt(self) -> ParamStmt:
    dram = self.dram()
    haoda_type = self.type_()
    attrs = []
    while self.accept(','):
      if self.accept('dup'):
        attrs.append(ParamAttr(dup=self.integer(), partitioning=None))
        continue
      self.expect('partition')
      strategy = self.ident()
      factor = dim = None
      if strategy == 'cyclic':
        self.expect('factor')
        self.expect('=')
        factor = self.integer()
      elif strategy != 'complete':
        raise SyntaxError_('unknown partition strategy %s' % strategy)
      if self.accept('dim'):
        self.expect('=')
        dim = self.integer()
      attrs.append(
          ParamAttr(dup=None,
                    partitioning=Partitioning(strategy=strategy,
                                              factor=factor,
                                              dim=dim)))
    self.expect(':')
    name = self.ident()
    size = []
    while self.accept('['):
      size.append(self.integer())
      self.expect(']')
    return ParamStmt(haoda_type=haoda_type,
                     name=name,
                     attr=attrs,
                     size=size,
                     dram=dram)

  def compute_stmt(self, cls):
    dram = self.dram() if cls is OutputStmt else ()
    haoda_type = self.type_()
    self.expect(':')
    lets = []
    while True:
      # Let: [Type] ID '=' Expr ;  Ref: ID '(' ...
      if self.is_type() and self.peek().kind == 'id' and \
          self.peek(2).text == '=':
        let_type = self.type_()
        name = self.ident()
        self.expect('=')
        lets.append(ir.Let(haoda_type=let_type, name=name, expr=self.expr()))
      elif self.tok.kind == 'id' and self.peek().text == '=':
        name = self.ident()
        self.expect('=')
        lets.append(ir.Let(haoda_type=None, name=name, expr=self.expr()))
      else:
        break
    ref = self.ref()
    self.expect('=')
    expr = self.expr()
    kwargs = dict(haoda_type=haoda_type, let=lets, ref=ref, expr=expr)
    if cls is OutputStmt:
      kwargs['dram'] = dram
    return cls(**kwargs)

  def ref(self) -> ir.Ref:
    name = self.ident()
    self.expect('(')
    idx = [self.integer(signed=True)]
    while self.accept(','):
      idx.append(self.integer(signed=True))
    self.expect(')')
    lat = None
    if self.accept('~'):
      lat = self.integer()
    return ir.Ref(name=name, idx=idx, lat=lat)

  # -- expressions ----------------------------------------------------------
  _LEVELS = (
      (ir.Expr, ('||',)),
      (ir.LogicAnd, ('&&',)),
      (ir.BinaryOr, ('|',)),
      (ir.Xor, ('^',)),
      (ir.BinaryAnd, ('&',)),
      (ir.EqCmp, ('==', '!=')),
      (ir.LtCmp, ('<=', '>=', '<', '>')),
      (ir.AddSub, ('+', '-')),
      (ir.MulDiv, ('*', '/', '%')),
  )

  def expr(self, level: int = 0) -> ir.Node:
    if level == len(self._LEVELS):
      return self.unary()
    cls, operators = self._LEVELS[level]
    operands = [self.expr(level + 1)]
    ops = []
    while self.tok.kind == 'op' and self.tok.text in operators:
      ops.append(self.tok.text)
      self.i += 1
      operands.append(self.expr(level + 1))
    if not ops:
      return operands[0]
    return cls(operand=operands, operator=ops)

  def unary(self) -> ir.Node:
    ops = []
    while self.tok.kind == 'op' and self.tok.text in ('+', '-', '~', '!'):
      ops.append(self.tok.text)
      self.i += 1
    operand = self.operand()
    if not ops:
      return operand
    return ir.Unary(operator=ops, operand=operand)

  def operand(self) -> ir.Node:
    tok = self.tok
    if tok.kind == 'num':
      self.i += 1
      return ir.Num(text=tok.text)
    if tok.kind == 'op' and tok.text == '(':
      self.i += 1
      inner = self.expr()
      self.expect(')')
      return ir.Operand(expr=inner)
    if tok.kind == 'id':
      nxt = self.peek()
      if nxt.kind == 'op' and nxt.text == '(':
        if self.is_type():
          haoda_type = self.type_()
          self.expect('(')
          inner = self.expr()
          self.expect(')')
          return ir.Cast(haoda_type=haoda_type, expr=inner)
        if tok.text in FUNC_NAMES:
          self.i += 2
          args = [self.expr()]
          while self.accept(','):
            args.append(self.expr())
          self.expect(')')
          return ir.Call(name=tok.text, arg=args)
        return self.ref()
      self.i += 1
      idx = []
      while self.at('['):
        self.i += 1
        idx.append(self.integer())
        self.expect(']')
      return ir.Var(name=tok.text, idx=idx)
    self.error('an operand')
    raise AssertionError  # unreachable


def parse(text: str) -> SodaProgram:
  """Parses SODA source text into a :class:`SodaProgram`."""
  return _Parser(text).program()


def parse_expr(text: str) -> ir.Node:
  parser = _Parser(text)
  node = parser.expr()
  if parser.tok.kind != 'eof':
    parser.error('end of expression')
  return node


def parse_file(path: str) -> SodaProgram:
  with open(path) as fp:
    return parse(fp.read())
