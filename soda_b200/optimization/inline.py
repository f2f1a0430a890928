"""``--inline`` and the always-on ``rebalance`` pass.

* ``inline``: a local tensor that is loaded exactly once in the whole program
  is substituted into its single use, shifted by the difference of indices
  (reference: src/soda/optimization/inline.py:13-80; behaviour pinned by
  src/tests/optimization/test_inline.py:24-113).
* ``rebalance``: a float ``+``-only sum with more than 32 terms is cut into
  groups of at most 32; all groups but the last become new ``cr_var_N`` locals
  stored at the statement's own index (the reference stores them at index 0,
  which is the same thing for every program whose long sum is stored at 0, and
  wrong otherwise; reference: src/soda/optimization/inline.py:170-262).  This
  changes floating-point association and therefore results, so the CUDA
  backend and the oracle both run it.
"""
import itertools
import logging
from typing import List, Tuple

from soda_b200 import grammar, ir, mutator, visitor

_logger = logging.getLogger(__name__)

REBALANCE_THRESHOLDS = {
    ir.Type('float'): 32,
}


def _loads_by_name(stencil):
  """{local name: set of (Ref, loading stmt)} over the whole program."""
  refs = {}
  for stmt in itertools.chain(stencil.local_stmts, stencil.output_stmts):
    loads = list(visitor.get_load_tuple(stmt.expr))
    for let in stmt.let:
      loads.extend(visitor.get_load_tuple(let))
    for ref in loads:
      if ref.name in stencil.input_names or ref.name == stmt.name or \
          ref.name in stencil.param_names:
        continue
      refs.setdefault(ref.name, []).append((ref, stmt))
  return refs


def inline(stencil):
  """Repeatedly inlines locals with exactly one load site."""
  while stencil.local_stmts:
    single = {
        name: uses[0]
        for name, uses in _loads_by_name(stencil).items()
        if len(uses) == 1
    }
    # never inline into a statement that is itself about to be inlined with a
    # stale body: take producers whose own loads are not candidates first
    choice = None
    for name in single:
      store_stmt = next(s for s in stencil.local_stmts if s.name == name)
      loaded = {ref.name for ref in visitor.get_load_set(store_stmt.expr)}
      for let in store_stmt.let:
        loaded |= {ref.name for ref in visitor.get_load_set(let)}
      if not loaded & set(single):
        choice = (name, store_stmt)
        break
    if choice is None:
      if not single:
        break
      name = next(iter(single))
      choice = (name, next(s for s in stencil.local_stmts if s.name == name))
    name, store_stmt = choice
    ref, load_stmt = single[name]

    # move the producer so that it is stored where the consumer loads it
    offset = tuple(a - b for a, b in zip(store_stmt.ref.idx, ref.idx))
    lets = tuple(mutator.shift(let, offset) for let in store_stmt.let)
    expr = mutator.shift(store_stmt.expr, offset)
    _logger.info('`%s` is referenced only once, replace with `%s`', ref, expr)
    if isinstance(expr, ir.BinaryOp) and not expr.singleton:
      expr = ir.Operand(expr=expr)

    def substitute(node, args):
      if isinstance(node, ir.Ref) and node == ref:
        return expr
      return node

    load_stmt.let = lets + tuple(
        let.visit(substitute) for let in load_stmt.let)
    load_stmt.expr = ir.unparenthesize(load_stmt.expr.visit(substitute))
    stencil.local_stmts.remove(store_stmt)
    stencil.invalidate()
  return stencil


def _split_sum(expr) -> List[ir.Node]:
  return list(expr.operand)


def rebalance(stencil):
  """Splits long float sums; see the module docstring."""
  changed = True
  while changed:
    changed = False
    for stmt in itertools.chain(stencil.local_stmts, stencil.output_stmts):
      threshold = REBALANCE_THRESHOLDS.get(stmt.haoda_type)
      if threshold is None:
        continue
      expr = ir.unparenthesize(stmt.expr)
      if not isinstance(expr, ir.AddSub) or set(expr.operator) != {'+'}:
        continue

      # (coefficient, sum) for `sum * coefficient` terms, else (None, term)
      terms: List[Tuple[object, ir.Node]] = []
      for operand in expr.operand:
        inner = ir.unwrap(operand) if isinstance(operand,
                                                 ir.Operand) else operand
        coeff = None
        body = operand
        if isinstance(inner, ir.MulDiv) and inner.operator == ('*',):
          lhs, rhs = (ir.unwrap(x) for x in inner.operand)
          if isinstance(lhs, ir.AddSub):
            coeff, body = inner.operand[1], lhs
          elif isinstance(rhs, ir.AddSub):
            coeff, body = inner.operand[0], rhs
        terms.append((coeff, body))

      def num_items(term) -> int:
        return 1 if term[0] is None else len(term[1].operand)

      terms.sort(key=num_items, reverse=True)  # stable
      groups: List[List[Tuple[object, ir.Node]]] = [[]]
      count = 0
      for term in terms:
        if count + num_items(term) > threshold:
          groups.append([])
          count = 0
        groups[-1].append(term)
        count += num_items(term)
      if len(groups) == 1:
        continue

      _logger.info('stmt %s has too many operations, breaking them into %d',
                   stmt.name, len(groups))
      new_exprs = []
      for group in groups:
        operands = []
        for coeff, body in group:
          if coeff is None:
            operands.append(body)
          else:
            operands.append(
                ir.MulDiv(operator=('*',),
                          operand=(ir.Operand(expr=body), coeff)))
        new_exprs.append(
            ir.AddSub(operator=('+',) * (len(operands) - 1),
                      operand=tuple(operands)))
      new_stmts = []
      for new_expr in new_exprs[:-1]:
        new_stmts.append(
            grammar.LocalStmt(ref=ir.Ref(name=stencil.new_cr_var(),
                                         lat=None,
                                         idx=stmt.ref.idx),
                              haoda_type=stmt.haoda_type,
                              expr=new_expr,
                              let=stmt.let,
                              stencil=stencil))
        stencil.local_stmts.append(new_stmts[-1])
      last = new_exprs[-1]
      stmt.expr = ir.AddSub(
          operator=last.operator + ('+',) * len(new_stmts),
          operand=last.operand + tuple(s.ref for s in new_stmts))
      stencil.invalidate()
      changed = True
      break
  return stencil
