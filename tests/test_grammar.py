"""Front-end tests, restating src/tests/test_grammar.py of the reference: the
program that uses every syntax feature must print back verbatim, and statement
nodes print like the reference's."""
import pytest

from soda_b200 import grammar, ir, util
from tests import common

FULL_PROGRAM = r'''
border: ignore
burst width: 512
cluster: none
iterate: 2
kernel: name
unroll factor: 1
input dram 0 float: bbb
input dram 1 uint6: a(233, *)
param int8: p0
param int9, dup 3: p1[23]
param int10, partition complete: p2[23]
param int11, partition complete dim=1: p2[23]
param int12, partition cyclic factor=23: p3[233]
param int13, partition cyclic factor=23 dim=2: p4[233][233]
param int14, partition complete, dup 3: p5[23]
local int27:
  int32 l = int32(a(0, 0) ~1 + b(1, 0))
  int32 g = int32(a(0, 0) ~1 + p0 + p1[1][3])
  c(0, 0) ~3 = +-+-l * --+~l
output dram 2 double:
  float18_3 l = float18_3(c(0, 1) ~5) + a(1, 0)
  d(0, 0) = sqrt(float15(l <= (l / 2)))
output dram 3 double:
  float18_3 l = float18_3(c(0, 1) ~5) + a(1, 0)
  e(0, 0) = float15(l + (l / 2))
'''.strip('\n')


def test_syntax_roundtrip():
  program = grammar.parse(FULL_PROGRAM)
  assert str(program) == FULL_PROGRAM
  assert program.dim == 2
  assert program.tile_size == (233, 0)


def _nodes():
  int8 = ir.Type('int8')
  ref = ir.Ref(name='foo', idx=(0, 23), lat=None)
  expr = ir.Expr(operand=(ir.Ref(name='bar', idx=(233, 42), lat=None),),
                 operator=())
  let = ir.Let(haoda_type=int8, name='foo_l',
               expr=ir.Expr(operand=(ir.Ref(name='bar_l', idx=(42, 2333),
                                            lat=None),), operator=()))
  let2 = ir.Let(haoda_type=int8, name='foo_l2',
                expr=ir.Expr(operand=(ir.Ref(name='bar_l2', idx=(0, 42),
                                             lat=None),), operator=()))
  return int8, ref, expr, let, let2


def test_input_stmt_str():
  int8 = ir.Type('int8')
  make = lambda tile: grammar.InputStmt(haoda_type=int8, name='foo',
                                        tile_size=tile, dram=())
  assert str(make([])) == 'input dram 0 int8: foo'
  assert str(make([23])) == 'input dram 0 int8: foo(23, *)'
  assert str(make([23, 233])) == 'input dram 0 int8: foo(23, 233, *)'


@pytest.mark.parametrize('cls,prefix', [(grammar.LocalStmt, 'local'),
                                        (grammar.OutputStmt, 'output dram 0')])
def test_compute_stmt_str(cls, prefix):
  int8, ref, expr, let, let2 = _nodes()
  kwargs = dict(haoda_type=int8, ref=ref, expr=expr)
  if cls is grammar.OutputStmt:
    kwargs['dram'] = ()
  assert str(cls(let=[], **kwargs)) == \
      prefix + ' int8: foo(0, 23) = bar(233, 42)'
  assert str(cls(let=[let], **kwargs)) == (
      prefix + ' int8:\n  int8 foo_l = bar_l(42, 2333)\n'
      '  foo(0, 23) = bar(233, 42)')
  assert str(cls(let=[let, let2], **kwargs)) == (
      prefix + ' int8:\n  int8 foo_l = bar_l(42, 2333)\n'
      '  int8 foo_l2 = bar_l2(0, 42)\n  foo(0, 23) = bar(233, 42)')


@pytest.mark.parametrize('name', common.PROGRAMS)
def test_shipped_programs_roundtrip(name):
  """Header directives may come in any order (textX unordered group,
  src/soda/grammar.py:15-30): parse -> print -> parse is a fixed point."""
  program = grammar.parse(common.source(name))
  again = grammar.parse(str(program))
  assert str(program) == str(again)
  assert program.app_name == name


def test_syntax_errors_are_semantic_errors():
  with pytest.raises(util.SemanticError):
    grammar.parse('kernel: x\nburst width: 64\nunroll factor: 1\n'
                  'input float: a(32, *)\noutput float: b(0, 0) = a(0, 0) +\n'
                  'iterate: 1')
  with pytest.raises(util.SemanticError):
    grammar.parse('kernel: x\nburst width: 64\nunroll factor: 1\niterate: 1\n'
                  'input float: a(32, *)')  # no output
  with pytest.raises(util.SemanticError):
    grammar.parse('kernel: x\nkernel: y\nburst width: 64\nunroll factor: 1\n'
                  'iterate: 1\ninput float: a(32, *)\n'
                  'output float: b(0, 0) = a(0, 0)')


def test_literal_types_follow_cxx():
  lit = lambda text: ir.Num(text=text).literal_type
  assert lit('3') == 'int32'
  assert lit('65535') == 'int32'
  assert lit('4294967295') == 'int64'
  assert lit('0xffffffff') == 'uint32'
  assert lit('7u') == 'uint32'
  assert lit('0.2f') == 'float'
  assert lit('.125f') == 'float'
  assert lit('2.0') == 'double'
  assert lit('1e3') == 'double'


def test_usual_arithmetic_conversions():
  t = ir.Type
  assert ir.common_type(t('uint16'), t('uint16')) == 'int32'
  assert ir.common_type(t('int16'), t('uint16')) == 'int32'
  assert ir.common_type(t('uint32'), t('int32')) == 'uint32'
  assert ir.common_type(t('uint32'), t('int64')) == 'int64'
  assert ir.common_type(t('int32'), t('float')) == 'float'
  assert ir.common_type(t('float'), t('double')) == 'double'
