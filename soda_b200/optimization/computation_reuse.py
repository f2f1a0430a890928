"""Computation reuse (``--computation-reuse``): share partial sums of a
reduction between neighbouring cells.

Model, as in the reference (reference:
src/soda/optimization/computation_reuse.py:1751-1861, ``Expression``): a pure
``+`` or pure ``*`` reduction is eligible when every operand loads exactly one
tensor index.  An operand is then the pair (relative attribute = that index,
absolute attribute = the operand normalised to index 0).  A *schedule* is a
binary tree over the operands; two sub-trees that are equal up to translation
compute the same values at shifted positions, so only one of them has to be
computed: it becomes a new local tensor ``cr_var_N`` and every occurrence
becomes a load of it (reference: ``CommSchedule.get_ir_node_with_cr``,
:755-868; statement creation :228-247).  The cost of a schedule is
``(num_ops, total_distance)``: distinct sub-trees, then the buffer span of the
shared ones (reference: ``CommSchedule.cost``, :404-410).

The search is this repo's own: a beam search over "merge the most frequent
translation-equivalent pair" moves, with exhaustive search for small
reductions.  It reproduces the operation counts pinned by the reference's
tests (src/tests/optimization/test_computation_reuse.py:174-352; restated in
tests/test_computation_reuse.py).  The reference's ``soda-cr`` binary,
``GloreSchedules`` and the ILP for ``total_distance`` are not reproduced; the
span-based distance here is an upper bound of the reference's definition for
chains and is only used to break ties.

On the GPU the new ``cr_var_N`` statements are ordinary stages of the fused
DAG: their values live in the register windows / shuffles of the kernel
templates like any other local (BASELINE.json north star: "the soda-cr
computation-reuse schedule becomes reuse of shared partial sums").
"""
import collections
import itertools
import json
import logging
import os
import subprocess
from typing import Dict, FrozenSet, Iterator, List, Optional, Sequence, Tuple

from soda_b200 import ir, mutator, util, visitor

_logger = logging.getLogger(__name__)

Index = Tuple[int, ...]
Leaf = Tuple[Index, int]  # (relative index, absolute attribute tag)

METHODS = ('yes', 'greedy', 'optimal', 'glore', 'beam', 'built-in',
           'built-in:greedy', 'built-in:optimal')


NATIVE_SOURCE = os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'csrc',
    'soda_cr', 'soda_cr.cpp')
NATIVE_BINARY = os.path.join(os.path.dirname(NATIVE_SOURCE), 'soda-cr')


def build_native(force: bool = False) -> Optional[str]:
  """Compiles the native scheduler (csrc/soda_cr/soda_cr.cpp) with g++."""
  if os.path.exists(NATIVE_BINARY) and not force and \
      os.path.getmtime(NATIVE_BINARY) >= os.path.getmtime(NATIVE_SOURCE):
    return NATIVE_BINARY
  tmp = '%s.%d.tmp' % (NATIVE_BINARY, os.getpid())
  result = subprocess.run(
      ['g++', '-O2', '-std=c++17', '-o', tmp, NATIVE_SOURCE],
      capture_output=True, text=True)
  if result.returncode != 0:
    _logger.warning('cannot build soda-cr: %s', result.stderr)
    return None
  os.replace(tmp, NATIVE_BINARY)
  return NATIVE_BINARY


class Linearizer:
  """Maps index tuples to integers such that differences of integers identify
  differences of tuples: dimension d gets ``2 * extent_d - 1`` slots (or the
  tile size, for all but the last dimension).  Same contract as the
  reference's (reference :72-154, pinned by
  src/tests/optimization/test_computation_reuse.py:33-42)."""

  def __init__(self, rattrs: Sequence[Sequence[int]],
               tile_size: Sequence[int] = ()):
    num_dim = len(rattrs[0])
    self.maxs = [max(r[d] for r in rattrs) for d in range(num_dim)]
    self.mins = [min(r[d] for r in rattrs) for d in range(num_dim)]
    if tile_size:
      self.sizes = tuple(tile_size)[:-1] + (
          (self.maxs[-1] - self.mins[-1] + 1) * 2 - 1,)
    else:
      self.sizes = tuple((self.maxs[d] - self.mins[d] + 1) * 2 - 1
                         for d in range(num_dim))

  @property
  def num_dim(self) -> int:
    return len(self.maxs)

  @property
  def dims(self) -> Tuple[int, ...]:
    return tuple(range(self.num_dim))

  @property
  def weights(self) -> List[int]:
    weights = [1] * self.num_dim
    for d in self.dims[1:]:
      weights[d] = weights[d - 1] * self.sizes[d - 1]
    return weights

  def apply(self, rattr: Sequence[int]) -> int:
    return sum((value - low) * weight
               for value, weight, low in zip(rattr, self.weights, self.mins))

  def restore(self, rattr: int) -> Tuple[int, ...]:
    restored = []
    for d in reversed(self.dims):
      value = rattr // self.weights[d]
      rattr -= value * self.weights[d]
      restored.append(self.mins[d] + value)
    return tuple(reversed(restored))

  def __call__(self, rattr):
    if isinstance(rattr, int):
      return self.restore(rattr)
    return self.apply(rattr)


def range_from_middle(n: int) -> Iterator[int]:
  """0..n-1 from the middle outwards (reference :157-174)."""
  middle = n // 2
  if n % 2:
    yield middle
    for shift in range(1, middle + 1):
      yield middle - shift
      yield middle + shift
  else:
    for shift in range(middle):
      yield middle - shift - 1
      yield middle + shift


def _sub(a: Index, b: Index) -> Index:
  return tuple(x - y for x, y in zip(a, b))


def _add(a: Index, b: Index) -> Index:
  return tuple(x + y for x, y in zip(a, b))


def _order(idx: Index):
  """Stream order: last dimension most significant
  (reference :1799-1802 sorts operands by reversed index)."""
  return tuple(reversed(idx))


class Pattern:
  """A sub-tree up to translation: the set of its leaves relative to its base
  (the position of its left-most leaf in stream order)."""
  __slots__ = ('leaves', 'left', 'right', 'distance', '_hash')

  def __init__(self, leaves: FrozenSet[Leaf], left=None, right=None,
               distance: Optional[Index] = None):
    self.leaves = leaves
    self.left = left        # Pattern or None for a leaf
    self.right = right
    self.distance = distance  # base(right) - base(left)
    self._hash = hash(leaves)

  def __hash__(self):
    return self._hash

  def __eq__(self, other):
    return isinstance(other, Pattern) and self.leaves == other.leaves

  @property
  def is_leaf(self) -> bool:
    return self.left is None

  @property
  def size(self) -> int:
    return len(self.leaves)

  def subtrees(self):
    """All non-leaf sub-trees, children first."""
    if self.is_leaf:
      return
    yield from self.left.subtrees()
    yield from self.right.subtrees()
    yield self

  @property
  def num_ops(self) -> int:
    return len(set(self.subtrees()))

  def __str__(self):
    if self.is_leaf:
      (_, tag), = self.leaves
      return str(tag)
    return '(%s==%s=>%s)' % (self.left, self.distance, self.right)


def leaf_pattern(tag: int, dim: int) -> Pattern:
  return Pattern(frozenset([((0,) * dim, tag)]))


def merge(left: Pattern, right: Pattern, distance: Index) -> Pattern:
  """left at base 0 combined with right at base ``distance``."""
  leaves = set(left.leaves)
  leaves.update((_add(idx, distance), tag) for idx, tag in right.leaves)
  return Pattern(frozenset(leaves), left, right, distance)


Item = Tuple[Index, Pattern]  # (base position, pattern)


def _pair_key(a: Item, b: Item):
  """Orders the pair in stream order and returns (left, right, distance)."""
  if (_order(a[0]), a[1]._hash) > (_order(b[0]), b[1]._hash):
    a, b = b, a
  return a, b, _sub(b[0], a[0])


def _disjoint_occurrences(items: Sequence[Item], left: Pattern,
                          right: Pattern, distance: Index):
  """Greedy maximal set of non-overlapping (i, j) with items[i] = left at p and
  items[j] = right at p + distance."""
  position = {}
  for i, (base, pattern) in enumerate(items):
    position.setdefault((base, pattern), []).append(i)
  used = set()
  found = []
  for i in sorted(range(len(items)), key=lambda k: _order(items[k][0])):
    if i in used or items[i][1] != left:
      continue
    target = (_add(items[i][0], distance), right)
    for j in position.get(target, ()):
      if j != i and j not in used:
        used.update((i, j))
        found.append((i, j))
        break
  return found


def _candidate_moves(items: Sequence[Item], limit: int):
  """Pair patterns that occur at least twice without overlap, best first."""
  counts = collections.Counter()
  for a, b in itertools.combinations(items, 2):
    left, right, distance = _pair_key(a, b)
    counts[(left[1], right[1], distance)] += 1
  moves = []
  ranked = sorted(counts.items(),
                  key=lambda kv: (-kv[1], sum(abs(x) for x in kv[0][2]),
                                  _order(kv[0][2]), kv[0][0]._hash,
                                  kv[0][1]._hash))
  best = 0
  for (left, right, distance), count in ranked:
    if count < 2 or (len(moves) >= limit and count < best):
      break
    if len(moves) >= 4 * limit:
      break
    occurrences = _disjoint_occurrences(items, left, right, distance)
    if len(occurrences) >= 2:
      best = max(best, len(occurrences))
      span = sum(abs(x) for x in distance)
      moves.append((-len(occurrences), span, _order(distance), left._hash,
                    right._hash, left, right, distance, occurrences))
  moves.sort(key=lambda m: m[:5])
  chosen = moves[:limit]
  # like the reference's greedy search (:1262-1276), also follow reuse along a
  # single dimension, highest dimension first: it keeps grids regular
  dim = len(items[0][0])
  for axis in reversed(range(dim)):
    aligned = [
        (key, count) for key, count in ranked if count >= 2 and
        key[2][axis] != 0 and all(key[2][d] == 0 for d in range(dim)
                                  if d != axis)
    ]
    extra = []
    for (left, right, distance), _ in aligned[:2 * limit]:
      occurrences = _disjoint_occurrences(items, left, right, distance)
      if len(occurrences) >= 2:
        extra.append((-len(occurrences), sum(abs(x) for x in distance),
                      _order(distance), left._hash, right._hash, left, right,
                      distance, occurrences))
    if extra:
      extra.sort(key=lambda m: m[:5])
      aligned_keys = {m[5:8] for m in extra[:2]}
      for move in extra[:2]:
        if all(move[5:8] != other[5:8] for other in chosen):
          chosen.append(move)
      return [m + (m[5:8] in aligned_keys,) for m in chosen]
  return [m + (False,) for m in chosen]


def _apply(items: Sequence[Item], left: Pattern, right: Pattern,
           distance: Index, occurrences) -> List[Item]:
  merged = merge(left, right, distance)
  consumed = set()
  result = []
  for i, j in occurrences:
    consumed.update((i, j))
    result.append((items[i][0], merged))
  result.extend(item for k, item in enumerate(items) if k not in consumed)
  return result


def _finish(items: Sequence[Item]) -> Pattern:
  """No more sharing possible: chain what is left in stream order."""
  ordered = sorted(items, key=lambda item: (_order(item[0]), item[1]._hash))
  base, tree = ordered[0]
  for position, pattern in ordered[1:]:
    tree = merge(tree, pattern, _sub(position, base))
  return tree


def total_distance(tree: Pattern) -> int:
  """Sum over shared sub-trees of the stream span between their first and last
  use (tie-breaker; see the module docstring)."""
  uses: Dict[Pattern, List[Index]] = collections.defaultdict(list)

  def walk(pattern: Pattern, base: Index):
    if pattern.is_leaf:
      return
    uses[pattern].append(base)
    if len(uses[pattern]) == 1:
      walk(pattern.left, base)
      walk(pattern.right, _add(base, pattern.distance))

  walk(tree, (0,) * len(next(iter(tree.leaves))[0]))
  total = 0
  for pattern, bases in uses.items():
    if len(bases) > 1:
      keys = [_order(b) for b in bases]
      lo, hi = min(keys), max(keys)
      total += sum(abs(a - b) for a, b in zip(hi, lo))
  return total


def find_schedule_native(leaves: Sequence[Leaf],
                         flag: Optional[str] = None) -> Optional[Pattern]:
  """Runs the native scheduler (csrc/soda_cr) over the reference's JSON
  contract (reference :1692-1743) and rebuilds the schedule as a Pattern tree.
  Returns None if the binary is unavailable or its answer does not cover the
  operands exactly."""
  if not os.path.exists(NATIVE_BINARY):
    return None
  linearizer = Linearizer([idx for idx, _ in leaves])
  request = {
      'rattrs': [linearizer.apply(idx) for idx, _ in leaves],
      'aattrs': [tag for _, tag in leaves],
      'linearizer': {'mins': linearizer.mins, 'maxs': linearizer.maxs,
                     'sizes': list(linearizer.sizes)},
  }
  try:
    result = subprocess.run([NATIVE_BINARY] + ([flag] if flag else []),
                            input=json.dumps(request), capture_output=True,
                            text=True, check=True, timeout=300)
    answer = json.loads(result.stdout)
  except (OSError, subprocess.SubprocessError, ValueError) as e:
    _logger.warning('soda-cr failed: %s', e)
    return None
  dim = len(leaves[0][0])

  def build(node, offset: int) -> Pattern:
    if not isinstance(node, dict):
      return leaf_pattern(int(node), dim)
    left = build(node['left'], offset)
    right_offset = offset + int(node['distance'])
    right = build(node['right'], right_offset)
    return merge(left, right, _sub(linearizer.restore(right_offset),
                                   linearizer.restore(offset)))

  tree = build(answer, min(request['rattrs']))
  base = min((idx for idx, _ in leaves), key=_order)
  expected = frozenset((_sub(idx, base), tag) for idx, tag in leaves)
  if tree.leaves != expected:
    _logger.warning('soda-cr returned a schedule that does not cover the '
                    'operands; falling back to the built-in search')
    return None
  return tree


def find_schedule(leaves: Sequence[Leaf], beam_width: int = 6,
                  branch: int = 4) -> Pattern:
  """Best schedule found for the reduction over ``leaves`` by the built-in
  (Python) beam search."""
  dim = len(leaves[0][0])
  start: List[Item] = [(idx, leaf_pattern(tag, dim)) for idx, tag in leaves]
  if len(leaves) <= 7:
    beam_width, branch = 64, 16
  elif len(leaves) > 128:
    beam_width, branch = 1, 1   # greedy: the reference also prunes harder
  elif len(leaves) > 48:        # as reductions grow (:1722-1731)
    beam_width, branch = 2, 2
  # a state is (items, regular): `regular` marks the lineage that only ever
  # followed single-dimension reuse; two such states are always kept so that
  # the regular decomposition of a grid cannot be crowded out of the beam
  frontier = [(start, True)]
  finished: List[Pattern] = []
  seen = set()

  def rank(state):
    items = state[0]
    merged = {p for _, p in items if not p.is_leaf}
    ops_so_far = len({q for p in merged for q in p.subtrees()})
    return (ops_so_far + len(items) - 1, len(items), sorted(
        (_order(b), p._hash) for b, p in items))

  while frontier:
    next_frontier = []
    for items, regular in frontier:
      moves = _candidate_moves(items, branch)
      if not moves:
        finished.append(_finish(items))
        continue
      for move in moves:
        left, right, distance, occurrences, aligned = move[5:]
        new_items = _apply(items, left, right, distance, occurrences)
        key = frozenset(collections.Counter(new_items).items())
        if key in seen:
          continue
        seen.add(key)
        next_frontier.append((new_items, regular and aligned))
    next_frontier.sort(key=rank)
    keep = next_frontier[:beam_width]
    protected = [s for s in next_frontier[beam_width:] if s[1]][:2]
    frontier = keep + protected
  best = min(finished, key=lambda t: (t.num_ops, total_distance(t), str(t)))
  return best


# ---------------------------------------------------------------------------
# IR side
# ---------------------------------------------------------------------------


class CannotHandle(Exception):
  pass


class Expression:
  """A reduction that is eligible for computation reuse
  (reference :1776-1838)."""

  def __init__(self, node: ir.Node, dim: int):
    reduction = ir.to_reduction(node)
    if reduction is None:
      raise CannotHandle(type(node).__name__)
    self.operator = reduction[0]
    operands = []
    for operand in reduction[1]:
      loads = visitor.get_load_set(operand)
      if len(loads) > 1:
        raise CannotHandle('multi-index operand %s' % operand)
      if not loads:
        raise CannotHandle('const operand %s' % operand)
      operands.append(operand)
    if len(operands) < 3:
      raise CannotHandle('nothing to share among %d operands' % len(operands))
    operands.sort(key=lambda x: _order(visitor.get_load_set(x)[0].idx))
    self.aattr_table: List[ir.Node] = []
    tags: Dict[str, int] = {}
    self.leaves: List[Leaf] = []
    for operand in operands:
      idx = visitor.get_load_set(operand)[0].idx
      normalised = mutator.shift(operand, idx)
      key = str(normalised)
      if key not in tags:
        tags[key] = len(self.aattr_table)
        self.aattr_table.append(normalised)
      self.leaves.append((idx, tags[key]))
    if len(set(self.leaves)) != len(self.leaves):
      raise CannotHandle('repeated operand')
    self.dim = dim


def _build_ir(expression: Expression, tree: Pattern, stencil,
              shared_refs: Dict[Pattern, str], new_stmts: list,
              stmt) -> ir.Node:
  """IR of ``tree`` placed at its own base; shared sub-trees become loads of
  cr_var locals (created on first use)."""
  from soda_b200 import grammar
  shared = {
      p for p, n in collections.Counter(_all_subtree_uses(tree)).items()
      if n > 1
  }

  def operand_ir(pattern: Pattern, base: Index) -> ir.Node:
    if pattern.is_leaf:
      (_, tag), = pattern.leaves
      return mutator.shift(expression.aattr_table[tag], base,
                           op=lambda a, b: a + b)
    if pattern in shared:
      if pattern not in shared_refs:
        name = stencil.new_cr_var()
        shared_refs[pattern] = name
        zero = (0,) * expression.dim
        body = combine(pattern, zero)
        new_stmts.append(
            grammar.LocalStmt(ref=ir.Ref(name=name, idx=zero, lat=None),
                              haoda_type=None,
                              expr=body,
                              let=stmt.let,
                              stencil=stencil))
      return ir.Ref(name=shared_refs[pattern], idx=base, lat=None)
    return ir.Operand(expr=combine(pattern, base))

  def combine(pattern: Pattern, base: Index) -> ir.Node:
    left = operand_ir(pattern.left, base)
    right = operand_ir(pattern.right, _add(base, pattern.distance))
    return ir.from_reduction(expression.operator, (left, right))

  root_base = min((idx for idx, _ in expression.leaves), key=_order)
  # the root is used once: never a cr_var itself
  return combine(tree, root_base)


def _all_subtree_uses(tree: Pattern):
  """Every non-leaf sub-tree occurrence, counting the inside of a shared
  sub-tree only once (it is computed once)."""
  seen = set()

  def walk(pattern: Pattern):
    if pattern.is_leaf:
      return
    yield pattern
    if pattern in seen:
      return
    seen.add(pattern)
    yield from walk(pattern.left)
    yield from walk(pattern.right)

  yield from walk(tree)


def computation_reuse(stencil):
  """Rewrites eligible reductions of ``stencil`` in place
  (reference :189-260)."""
  method = stencil.optimizations.get('computation-reuse')
  if method is None or method == 'no':
    return stencil
  if method not in METHODS:
    raise util.SemanticError('unknown computation reuse method `%s`' % method)
  import itertools as _it
  new_stmts: list = []

  def rewrite(node, stmt):
    def visit(obj, args):
      if not isinstance(obj, (ir.AddSub, ir.MulDiv)):
        return obj
      try:
        expression = Expression(obj, stencil.dim)
      except CannotHandle:
        return obj
      # like the reference (:1840-1856): the external tool when it exists and
      # the method does not ask for the built-in search, else the built-in one
      tree = None
      if not method.startswith('built-in'):
        flag = {'greedy': '--greedy', 'optimal': '--brute-force',
                'beam': '--beam'}.get(method)
        tree = find_schedule_native(expression.leaves, flag)
      if tree is None:
        tree = find_schedule(expression.leaves)
      if tree.num_ops >= len(expression.leaves) - 1:
        return obj  # nothing gained
      _logger.info('%s: %d operations instead of %d: %s', stmt.name,
                   tree.num_ops, len(expression.leaves) - 1, tree)
      shared_refs: Dict[Pattern, str] = {}
      replacement = _build_ir(expression, tree, stencil, shared_refs,
                              new_stmts, stmt)
      return ir.Operand(expr=replacement)

    return node.visit(visit)

  for stmt in _it.chain(list(stencil.local_stmts), stencil.output_stmts):
    stmt.expr = ir.unparenthesize(rewrite(stmt.expr, stmt))
    stmt.let = tuple(rewrite(let, stmt) for let in stmt.let)

  if new_stmts:
    # type the new locals (their expressions only load existing tensors or
    # earlier cr_vars)
    for new_stmt in new_stmts:
      stencil.local_stmts.append(new_stmt)
      stencil.invalidate()
    stencil.invalidate()
    pending = list(new_stmts)
    while pending:
      progress = False
      for new_stmt in list(pending):
        names = {r.name for r in visitor.get_load_set(new_stmt.expr)}
        if names & {s.name for s in pending if s is not new_stmt}:
          continue
        table = dict(stencil.symbol_table)
        for other in new_stmts:
          if other.haoda_type is None:
            table.pop(other.name, None)
        new_stmt.haoda_type = ir.propagate_type(new_stmt.expr,
                                                new_stmt.symbol_table
                                                if new_stmt.let else
                                                table).haoda_type
        stencil.invalidate()
        pending.remove(new_stmt)
        progress = True
      if not progress:
        raise util.InternalError('cyclic computation-reuse variables')
  return stencil
