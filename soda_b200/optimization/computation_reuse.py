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


def _candidate_moves(items: Sequence[Item], limit: int,
                     axis_order: Optional[Sequence[int]] = None):
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
  for axis in (axis_order if axis_order is not None else
               tuple(reversed(range(dim)))):
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


def reuse_dependencies(tree: Pattern, weights: Sequence[int]):
  """Dependency tables between the reused variables of a schedule, as the
  reference computes them (reference :413-541, ``_calc_dependency``): variable
  0 is the input, 1 the whole schedule, 2.. the sub-trees that occur more than
  once (translation-equivalent), in pre-order of first occurrence.  A variable
  that a single other variable reads at a single offset is inlined into it.

  Returns ``(dependers, dependees)``: ``dependers[src]`` is the ordered set of
  variables that read ``src``; ``dependees[dst][src]`` the (first, last)
  linearised offset at which ``dst`` reads ``src``."""
  def linear(distance: Index) -> int:
    return sum(d * w for d, w in zip(distance, weights))

  def occurrences(pattern: Pattern):
    if pattern.is_leaf:
      return
    yield pattern
    yield from occurrences(pattern.left)
    yield from occurrences(pattern.right)

  counts: Dict[Pattern, int] = collections.OrderedDict()
  for pattern in occurrences(tree):
    counts[pattern] = counts.get(pattern, 0) + 1
  var_of: Dict[Pattern, int] = collections.OrderedDict([(tree, 1)])
  table = {1: tree}
  for pattern, count in counts.items():
    if count > 1 and pattern not in var_of:
      var_of[pattern] = len(var_of) + 1
      table[len(var_of)] = pattern

  def accesses(pattern: Pattern, offset: Optional[int] = None):
    vid = var_of.get(pattern)
    if vid is not None and offset is not None:
      yield offset, vid
      return
    offset = offset or 0
    for child, base in ((pattern.left, offset),
                        (pattern.right, offset + linear(pattern.distance))):
      if child.is_leaf:
        yield base, 0
      else:
        yield from accesses(child, base)

  dependers: Dict[int, Dict[int, None]] = collections.OrderedDict()
  dependees: Dict[int, Dict[int, Tuple[int, int]]] = collections.OrderedDict()
  queue = collections.deque([tree])
  processed = {0}
  while queue:
    pattern = queue.popleft()
    dst = var_of[pattern]
    processed.add(dst)
    for offset, src in accesses(pattern):
      dependers.setdefault(src, collections.OrderedDict()).setdefault(dst)
      lo, hi = dependees.setdefault(
          dst, collections.OrderedDict()).setdefault(src, (offset, offset))
      dependees[dst][src] = (min(lo, offset), max(hi, offset))
      if src not in processed and table[src] not in queue:
        queue.append(table[src])

  def find_inline():
    for src, dsts in dependers.items():
      if len(dsts) == 1:
        dst = next(iter(dsts))
        lo, hi = dependees[dst][src]
        if lo == hi:
          return src, dst
    return None

  while True:
    found = find_inline()
    if found is None:
      break
    src, dst = found
    offset = dependees[dst][src][0]
    for inner, (lo, hi) in dependees.get(src, {}).items():
      new = (lo + offset, hi + offset)
      old = dependees[dst].get(inner, new)
      dependees[dst][inner] = (min(old[0], new[0]), max(old[1], new[1]))
    for inner in list(dependees.get(src, {})):
      dependers[inner][dst] = None
      del dependers[inner][src]
    del dependers[src]
    del dependees[dst][src]
    dependees.pop(src, None)
  return dependers, dependees


def reference_total_distance(tree: Pattern, linearizer: 'Linearizer') -> int:
  """The reference's ``CommSchedule.total_distance`` (reference :573-624): the
  sum over reused variables (the input included) of the distance between the
  offset at which a variable is first produced and the one at which it is last
  consumed, with the produce offsets chosen by an integer program.  The
  program has difference constraints only, so its LP relaxation (solved here
  with scipy's HiGHS; the reference uses pulp + CBC) is integral."""
  import numpy as np
  from scipy.optimize import linprog
  if tree.is_leaf:
    return 0
  dependers, dependees = reuse_dependencies(tree, linearizer.weights)
  sources = list(dependers)
  # unknowns: produce offset P_v of every source but the input (P_0 = P_1 = 0),
  # consume offset C_v of every source
  p_index = {v: i for i, v in enumerate(v for v in sources if v not in (0, 1))}
  c_index = {v: len(p_index) + i for i, v in enumerate(sources)}
  n = len(p_index) + len(c_index)
  cost = np.zeros(n)
  for v in sources:
    cost[c_index[v]] += 1
    if v in p_index:
      cost[p_index[v]] -= 1
  rows, rhs = [], []
  for src, dsts in dependers.items():
    for dst in dsts:
      lo, hi = dependees[dst][src]
      # P_src <= lo + P_dst
      row = np.zeros(n)
      if src in p_index:
        row[p_index[src]] += 1
      if dst in p_index:
        row[p_index[dst]] -= 1
      rows.append(row)
      rhs.append(lo)
      # C_src >= hi + P_dst
      row = np.zeros(n)
      row[c_index[src]] -= 1
      if dst in p_index:
        row[p_index[dst]] += 1
      rows.append(row)
      rhs.append(-hi)
  result = linprog(cost, A_ub=np.array(rows), b_ub=np.array(rhs),
                   bounds=[(None, None)] * n, method='highs')
  if result.status != 0:
    raise util.InternalError('offset program failed: %s' % result.message)
  return int(round(result.fun))


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

  base = min((idx for idx, _ in leaves), key=_order)
  expected = frozenset((_sub(idx, base), tag) for idx, tag in leaves)
  origin = min(request['rattrs'])
  candidates = [build(answer, origin)]
  # --optimal also returns every other tree with as few operations (up to a
  # cap): the tie is broken by total reuse distance, the second component of
  # the reference's schedule cost (:398-400)
  for other in answer.get('alternatives', ()):
    candidates.append(build(other, origin))
  if any(tree.leaves != expected for tree in candidates):
    _logger.warning('soda-cr returned a schedule that does not cover the '
                    'operands; falling back to the built-in search')
    return None
  if flag == '--optimal' and not answer.get('proven_optimal', False):
    _logger.warning('soda-cr --optimal stopped at its tree budget after %s '
                    'trees: best schedule found, optimality not proven',
                    answer.get('trees'))
  if len(candidates) == 1:
    return candidates[0]
  return min(candidates,
             key=lambda t: (t.num_ops, reference_total_distance(t, linearizer),
                            str(t)))


def find_schedule_exhaustive(leaves: Sequence[Leaf],
                             limit: int = 8) -> Optional[Pattern]:
  """Every binary tree over the operands (reference :983-1059,
  ``CommSchedules.generator``: the first operand stays left, every subset of
  the rest joins it), cost = (operations, total reuse distance).  Pure
  Python: only up to ``limit`` operands; the native soda-cr does more."""
  if len(leaves) > limit:
    return None
  dim = len(leaves[0][0])
  ordered = sorted(leaves, key=lambda leaf: (_order(leaf[0]), leaf[1]))
  linearizer = Linearizer([idx for idx, _ in leaves])

  def trees(indices: Tuple[int, ...]):
    if len(indices) == 1:
      yield leaf_pattern(ordered[indices[0]][1], dim)
      return
    rest = indices[1:]
    for size in range(len(rest)):
      for chosen in itertools.combinations(rest, size):
        left = (indices[0],) + chosen
        right = tuple(i for i in rest if i not in chosen)
        distance = _sub(ordered[right[0]][0], ordered[left[0]][0])
        for l in trees(left):
          for r in trees(right):
            yield merge(l, r, distance)

  best, best_cost = None, None
  for tree in trees(tuple(range(len(ordered)))):
    ops = tree.num_ops
    if best_cost is not None and ops > best_cost[0]:
      continue
    cost = (ops, reference_total_distance(tree, linearizer), str(tree))
    if best_cost is None or cost < best_cost:
      best, best_cost = tree, cost
  return best


def _linear_chain(items: Sequence[Item]) -> Item:
  """Right-deep chain over the items in stream order: i0 + (i1 + (i2 + ...))
  (reference :1503-1520, ``linear_schedule``).  Returns (base, pattern)."""
  ordered = sorted(items, key=lambda item: (_order(item[0]), item[1]._hash))
  base, tree = ordered[-1]
  for position, pattern in reversed(ordered[:-1]):
    tree = merge(pattern, tree, _sub(base, position))
    base = position
  return base, tree


def find_schedule_glore(leaves: Sequence[Leaf]) -> Pattern:
  """The GLORE heuristic as the reference applies it (reference :1523-1689,
  ``GloreSchedules``), once along dimension 0 and once along the diagonal:

  1. operands are grouped into lines along the direction;
  2. inside a line of more than three operands, the stride that pairs the most
     operands into translates of (operand at the line's end, operand one
     stride before it) is chosen, and every such pair becomes one shared
     operand;
  3. lines that then look alike (same stride, same pairing, same tags) are
     computed once, as a chain, and reused for every such line;
  4. what is left is chained in stream order.
  The direction with fewer operations wins."""
  dim = len(leaves[0][0])
  results = []
  for diagonal in (False, True):
    lines: Dict[Index, List[Leaf]] = collections.OrderedDict()
    for idx, tag in leaves:
      line_id = tuple(x - idx[0] for x in idx[1:]) if diagonal else idx[1:]
      lines.setdefault(line_id, []).append((idx, tag))
    by_key: Dict[tuple, List[List[Item]]] = collections.OrderedDict()
    for line_id, line in lines.items():
      line.sort(key=lambda leaf: _order(leaf[0]), reverse=True)
      dists = [line[0][0][0] - idx[0] for idx, _ in line]
      at = dict(zip(dists, line))
      chosen = None
      if len(line) > 3:
        for stride in range(dists[1], dists[-1]):
          if stride not in at:
            continue
          head = (at[0][1], at[stride][1])
          pending = list(dists)
          reused, plain, items = [], [], []
          while pending:
            d = pending.pop(0)
            if d + stride in pending and (at[d][1], at[d + stride][1]) == head:
              pending.remove(d + stride)
              reused.append(d)
              lower, upper = at[d + stride][0], at[d][0]
              pair = merge(leaf_pattern(at[stride][1], dim),
                           leaf_pattern(at[0][1], dim), _sub(upper, lower))
              items.append((lower, pair))
            else:
              plain.append(d)
              items.append((at[d][0], leaf_pattern(at[d][1], dim)))
          if reused and (chosen is None or
                         (len(reused), -stride) > (len(chosen[1]), -chosen[0])):
            chosen = (stride, tuple(reused), tuple(plain), items)
      if chosen is None:
        chosen = (0, (), tuple(dists),
                  [(idx, leaf_pattern(tag, dim)) for idx, tag in line])
      stride, reused, plain, items = chosen
      items.sort(key=lambda item: _order(item[0]))
      shape = tuple((_sub(pos, items[0][0]), pattern) for pos, pattern in items)
      by_key.setdefault((stride, reused, plain, shape), []).append(items)
    final: List[Item] = []
    for (stride, reused, plain, _), groups in by_key.items():
      if len(groups) > 1 and len(reused) + len(plain) > 1:
        for items in groups:  # one shared chain, used once per line
          final.append(_linear_chain(items))
      else:
        for items in groups:
          final.extend(items)
    results.append(_linear_chain(final)[1] if len(final) > 1 else final[0][1])
  return min(results, key=lambda tree: (tree.num_ops, str(tree)))


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
  # the regular decomposition of a grid cannot be crowded out of the beam.
  # The search runs once per preference order of the dimensions for that
  # single-dimension reuse: the orders reach the same operation count on
  # regular grids but very different reuse distances (an 11 x 11 box: 220
  # when rows are summed first, 374 when columns are).
  finished: List[Pattern] = []

  def rank(state):
    items = state[0]
    merged = {p for _, p in items if not p.is_leaf}
    ops_so_far = len({q for p in merged for q in p.subtrees()})
    return (ops_so_far + len(items) - 1, len(items), sorted(
        (_order(b), p._hash) for b, p in items))

  orders = [tuple(reversed(range(dim)))]
  if dim > 1:
    orders.append(tuple(range(dim)))
  for axis_order in orders:
    frontier = [(start, True)]
    seen = set()
    while frontier:
      next_frontier = []
      for items, regular in frontier:
        moves = _candidate_moves(items, branch, axis_order)
        if not moves:
          # no sharing left: chain the rest, left-deep and right-deep (same
          # operations, different reuse distances)
          finished.append(_finish(items))
          if len(items) > 2:
            finished.append(_linear_chain(items)[1])
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
  # cost = (operations, total reuse distance) as in the reference (:398-400);
  # the distance needs an LP per candidate, so only the candidates with the
  # fewest operations are measured
  fewest = min(t.num_ops for t in finished)
  linearizer = Linearizer([idx for idx, _ in leaves])
  # (Pattern equality is "same leaves": every complete tree is equal to every
  # other one, so the candidates are told apart by their text)
  finalists = {str(t): t for t in finished if t.num_ops == fewest}
  return min(finalists.values(),
             key=lambda t: (reference_total_distance(t, linearizer), str(t)))


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
        flag = {'greedy': '--greedy', 'optimal': '--optimal',
                'beam': '--beam'}.get(method)
        if method == 'glore':
          tree = find_schedule_glore(expression.leaves)
        elif method == 'optimal' and len(expression.leaves) > 16:
          raise util.SemanticError(
              '--computation-reuse=optimal enumerates every schedule: at most '
              '16 operands per reduction (this one has %d); use `yes`' %
              len(expression.leaves))
        else:
          tree = find_schedule_native(expression.leaves, flag)
      elif method.endswith('optimal'):
        tree = find_schedule_exhaustive(expression.leaves)
        if tree is None:
          raise util.SemanticError(
              '--computation-reuse=built-in:optimal handles at most 8 '
              'operands per reduction (this one has %d); `optimal` uses the '
              'native search' % len(expression.leaves))
      if tree is None:
        if method == 'optimal':
          tree = find_schedule_exhaustive(expression.leaves)
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
        new_stmt.haoda_type = ir.propagate_type(
            new_stmt.expr, new_stmt.symbol_table if new_stmt.let else table,
            getattr(stencil, 'float_math', False)).haoda_type
        stencil.invalidate()
        pending.remove(new_stmt)
        progress = True
      if not progress:
        raise util.InternalError('cyclic computation-reuse variables')
  return stencil
