"""Stencil IR and analyses: the API that stays.

``Stencil(**kwargs)`` takes the same keyword arguments ``sodac`` passes in the
reference (reference: src/soda/sodac.py:173-194, src/soda/core.py:52-76), runs
the same pass pipeline (computation reuse -> inline -> rebalance -> type
propagation, reference: src/soda/core.py:134-142) and exposes the same
attributes: ``tensors`` (the DAG unrolled over ``iterate``),
``chronological_tensors``, ``stencil_window``, ``stencil_distance``,
``reuse_buffers`` and friends.

Differences, all forced by the environment or by the target:

* no FPGA dataflow graph / module clustering is built (reference:
  src/soda/core.py:139,145-154) - the CUDA backend fuses everything;
* the produce/consume offsets the reference obtains from a CBC ILP
  (reference: src/soda/core.py:371-446; ``pulp`` is not installed) are computed
  by a longest-path (as-soon-as-possible) schedule, which is feasible for the
  ILP's constraints and optimal for chain-shaped programs;
* ``window_bounds`` gives the per-tensor bounding box of the dependency cone in
  closed form, so ``iterate: 256`` does not have to enumerate a 131 k-point
  window.
"""
import collections
import itertools
import logging
from typing import Dict, List, Sequence, Tuple

from soda_b200 import grammar, ir, util, visitor
from soda_b200 import tensor as soda_tensor
from soda_b200.optimization import computation_reuse as cr
from soda_b200.optimization import inline

_logger = logging.getLogger(__name__)


def _cached(func):
  """cached_property that stores into the instance ``__dict__`` under the
  function's name, so ``stencil.__dict__.pop(name)`` invalidates it (the
  reference relies on this, reference: src/soda/optimization/inline.py:70-72)."""
  name = func.__name__

  class _Descriptor:

    def __get__(self, obj, cls):
      if obj is None:
        return self
      value = obj.__dict__[name] = func(obj)
      return value

  _Descriptor.__doc__ = func.__doc__
  return _Descriptor()


class Stencil:
  """A validated, optimised SODA program.

  Attributes (same meaning as the reference's):
    iterate, border, preserve_border, cluster, burst_width, app_name,
    tile_size, unroll_factor, replication_factor, dim, param_stmts,
    input_stmts, local_stmts, output_stmts, optimizations.
  """

  def __init__(self, **kwargs):
    self.iterate = kwargs.pop('iterate')
    if self.iterate < 1:
      raise util.SemanticError('cannot iterate %d times' % self.iterate)
    self.border = kwargs.pop('border', None) or 'ignore'
    self.preserve_border = self.border == 'preserve'
    self.cluster = kwargs.pop('cluster', None) or 'none'
    self.burst_width = kwargs.pop('burst_width')
    self.app_name = kwargs.pop('app_name')
    self.tile_size = tuple(kwargs.pop('tile_size'))
    self.unroll_factor = kwargs.pop('unroll_factor')
    self.replication_factor = kwargs.pop('replication_factor', 1)
    self.dim = kwargs.pop('dim')
    self.param_stmts = list(kwargs.pop('param_stmts', ()) or ())
    self.input_stmts = list(kwargs.pop('input_stmts'))
    self.local_stmts = list(kwargs.pop('local_stmts', ()) or ())
    self.output_stmts = list(kwargs.pop('output_stmts'))
    self.optimizations = dict(kwargs.pop('optimizations', None) or {})
    # 'double' (default) or 'float': see ir.FLOAT_MATH_CALLS
    self.math_precision = kwargs.pop('math_precision', None) or 'double'
    if self.math_precision not in ('double', 'float'):
      raise util.SemanticError('math precision must be double or float')
    self.float_math = self.math_precision == 'float'

    self._override_dram(kwargs.pop('dram_in', None), self.input_stmts, '^',
                        'input')
    self._override_dram(kwargs.pop('dram_out', None), self.output_stmts, ',',
                        'output')

    if self.iterate > 1:
      if len(self.input_stmts) != len(self.output_stmts):
        raise util.SemanticError(
            'number of input tensors must be the same as output if iterate > 1 '
            'times, currently there are %d input(s) but %d output(s)' %
            (len(self.input_stmts), len(self.output_stmts)))
      if self.input_types != self.output_types:
        raise util.SemanticError(
            'input must have the same type(s) as output if iterate > 1 '
            'times, current input has type %s but output has type %s' %
            (util.lst2str(self.input_types), util.lst2str(self.output_types)))

    for stmt in itertools.chain(self.local_stmts, self.output_stmts):
      stmt.stencil = self
      if len(stmt.ref.idx) != self.dim:
        raise util.SemanticError(
            '`%s` is stored with %d indices in a %d-dimensional program' %
            (stmt.name, len(stmt.ref.idx), self.dim))

    self._check_references()

    self._cr_counter = 0
    cr.computation_reuse(self)
    if 'inline' in self.optimizations:
      inline.inline(self)
    inline.rebalance(self)

    for stmt in itertools.chain(self.local_stmts, self.output_stmts):
      stmt.propagate_type()

  # -- construction helpers -------------------------------------------------
  @staticmethod
  def _override_dram(spec, stmts, separator, kind) -> None:
    """``--dram-in`` / ``--dram-out`` overrides, parsed like the reference
    (reference: src/soda/core.py:78-106); banks are meaningless on a GPU."""
    if spec is None:
      return
    if ':' in spec:
      table = {stmt.name: stmt for stmt in stmts}
      for entry in spec.split(separator):
        name, banks = entry.split(':')
        if name not in table:
          raise util.SemanticError('no {} named `{}`'.format(kind, name))
        table[name].dram = tuple(map(int, banks.split('.')))
    else:
      for stmt in stmts:
        stmt.dram = tuple(map(int, spec.split('.')))

  def _check_references(self) -> None:
    known = set(self.input_names)
    params = set(self.param_names)
    all_names = set(self.input_names + self.local_names + self.output_names)
    for stmt in itertools.chain(self.local_stmts, self.output_stmts):
      for ref in visitor.get_load_tuple(stmt.expr) + tuple(
          r for let in stmt.let for r in visitor.get_load_tuple(let)):
        if ref.name in params:
          continue
        if ref.name not in all_names:
          raise util.SemanticError('`%s` loads undefined tensor `%s`' %
                                   (stmt.name, ref.name))
        if len(ref.idx) != self.dim:
          raise util.SemanticError(
              '`%s` is loaded with %d indices in a %d-dimensional program' %
              (ref.name, len(ref.idx), self.dim))
      known.add(stmt.name)

  def __str__(self) -> str:
    stmts = (self.input_stmts + self.param_stmts + self.local_stmts +
             self.output_stmts)
    return ('kernel: {0.app_name}\nburst width: {0.burst_width}\n'
            'iterate: {0.iterate}\nunroll factor: {0.unroll_factor}\n{1}\n'
            'border: {0.border}\ncluster: {0.cluster}').format(
                self, '\n'.join(map(str, stmts)))

  @property
  def kernel_name(self) -> str:
    return '%s_kernel' % self.app_name

  def new_cr_var(self) -> str:
    taken = {
        stmt.name
        for stmt in self.input_stmts + self.local_stmts + self.output_stmts
    }
    while True:
      var = 'cr_var_%d' % self._cr_counter
      self._cr_counter += 1
      if var not in taken:
        return var

  def invalidate(self) -> None:
    """Drops every cached analysis after statements were rewritten."""
    for name in ('symbol_table', 'local_names', 'local_types', 'stmt_table',
                 'tensors', 'chronological_tensors', 'window_bounds',
                 'reuse_buffers', 'all_points', 'next_fifo'):
      self.__dict__.pop(name, None)
    self.__dict__.pop('_stencil_window', None)
    self.__dict__.pop('_stencil_distance', None)

  # -- tables -----------------------------------------------------------------
  @_cached
  def stmt_table(self):
    return {
        stmt.name: stmt for stmt in self.input_stmts + self.local_stmts +
        self.output_stmts + self.param_stmts
    }

  @property
  def input_types(self):
    return tuple(stmt.haoda_type for stmt in self.input_stmts)

  @property
  def param_types(self):
    return tuple(stmt.haoda_type for stmt in self.param_stmts)

  @_cached
  def local_types(self):
    return tuple(stmt.haoda_type for stmt in self.local_stmts)

  @property
  def output_types(self):
    return tuple(stmt.haoda_type for stmt in self.output_stmts)

  @property
  def input_names(self):
    return tuple(stmt.name for stmt in self.input_stmts)

  @property
  def param_names(self):
    return tuple(stmt.name for stmt in self.param_stmts)

  @_cached
  def local_names(self):
    return tuple(stmt.name for stmt in self.local_stmts)

  @property
  def output_names(self):
    return tuple(stmt.name for stmt in self.output_stmts)

  @_cached
  def symbol_table(self) -> Dict[str, ir.Type]:
    """{tensor name: type}; let variables are per statement
    (``stmt.symbol_table``)."""
    table: Dict[str, ir.Type] = {}
    for stmt in itertools.chain(self.input_stmts, self.param_stmts,
                                self.local_stmts, self.output_stmts):
      if stmt.name in table:
        raise util.InputError('conflicting stmt name: %s' % stmt.name)
      table[stmt.name] = stmt.haoda_type
    return table

  @property
  def propagate_type(self):
    """Callable ``(node, stmt=None)`` typing ``node`` in this program's (or the
    statement's) scope (reference: src/soda/core.py:255-272)."""

    def propagate(node, stmt=None):
      table = self.symbol_table if stmt is None else stmt.symbol_table
      return ir.propagate_type(node, table, self.float_math)

    return propagate

  # -- the DAG unrolled over `iterate` -----------------------------------------
  def name_in_iter(self, name: str, iteration: int) -> str:
    """Tensor name of ``name`` in iteration ``iteration`` of the chain
    ``in -> locals -> in_iter1 -> ... -> out``
    (reference: src/soda/core.py:320-336)."""
    if name in self.input_names:
      return name + '_iter%d' % iteration if iteration > 0 else name
    if name in self.output_names:
      if iteration < self.iterate - 1:
        return (self.input_names[self.output_names.index(name)] +
                '_iter%d' % (iteration + 1))
      return name
    if name in self.local_names:
      return name + '_iter%d' % iteration if iteration > 0 else name
    if name in self.param_names:
      return name
    raise util.InternalError('unknown name: %s' % name)

  @_cached
  def tensors(self) -> Dict[str, soda_tensor.Tensor]:
    """Ordered {name: Tensor} over all ``iterate`` copies of the DAG."""
    tensor_map: Dict[str, soda_tensor.Tensor] = collections.OrderedDict()
    for stmt in self.input_stmts:
      tensor_map[stmt.name] = soda_tensor.Tensor(stmt, self.tile_size)

    symbol_table = self.symbol_table
    for iteration in range(self.iterate):

      def rename(obj, args):
        if isinstance(obj, ir.Ref):
          obj.haoda_type = symbol_table[obj.name]
          obj.name = self.name_in_iter(obj.name, iteration)  # pylint: disable=cell-var-from-loop
        return obj

      tensors = []
      for stmt in itertools.chain(self.local_stmts, self.output_stmts):
        tensor = soda_tensor.Tensor(stmt.visit(rename), self.tile_size)
        if tensor.name in tensor_map:
          raise util.InputError('conflicting tensor name: %s' % tensor.name)
        tensor_map[tensor.name] = tensor
        tensors.append(tensor)

      for tensor in tensors:
        for parent_name, ld_refs in visitor.get_load_dict(tensor).items():
          if parent_name in self.param_names:
            continue
          ld_refs = sorted(
              ld_refs, key=lambda ref: util.serialize(ref.idx, self.tile_size))
          parent = tensor_map[parent_name]
          parent.children[tensor.name] = tensor
          tensor.parents[parent_name] = parent
          tensor.ld_refs[parent_name] = ld_refs

    self._schedule(tensor_map)
    return tensor_map

  def _schedule(self, tensor_map) -> None:
    """Stream offsets at which each tensor is produced / may be dropped.

    Constraints are the reference ILP's (reference: src/soda/core.py:392-403):
    a tensor element is produced no earlier than the newest element it loads,
    and a parent element is kept until its oldest use.  The ASAP solution is
    used instead of the ILP optimum.
    """
    order = _toposort({t.name: set(t.parents) for t in tensor_map.values()})
    produced: Dict[str, int] = {}
    for name in order:
      tensor = tensor_map[name]
      earliest = 0
      for parent_name, offsets in tensor.ld_offsets.items():
        # produced[parent] <= produced[self] + st_offset - newest_access
        earliest = max(
            earliest,
            produced[parent_name] - (tensor.st_offset - max(offsets)))
      produced[name] = earliest if tensor.parents else 0
    base = min(produced[name] for name in self.input_names)
    for tensor in tensor_map.values():
      tensor.produce_offset = produced[tensor.name] - base
      tensor.consume_offset = tensor.produce_offset
      tensor.max_access = 0
    for parent in tensor_map.values():
      for child in parent.children.values():
        offsets = child.ld_offsets[parent.name]
        oldest = (child.st_offset - min(offsets) + child.produce_offset -
                  parent.produce_offset)
        parent.max_access = max(parent.max_access, oldest)
        parent.consume_offset = max(
            parent.consume_offset,
            child.produce_offset + child.st_offset - min(offsets))

  @_cached
  def chronological_tensors(self) -> List[soda_tensor.Tensor]:
    """Tensors in dependency order, names sorted inside each level."""
    return [
        self.tensors[name] for name in _toposort(
            {t.name: set(t.parents) for t in self.tensors.values()})
    ]

  @property
  def producer_tensors(self):
    return tuple(t for t in self.tensors.values() if t.is_producer())

  @property
  def consumer_tensors(self):
    return tuple(t for t in self.tensors.values() if t.is_consumer())

  # -- windows ------------------------------------------------------------------
  @_cached
  def window_bounds(self) -> Dict[str, Tuple[Tuple[int, ...], Tuple[int, ...]]]:
    """{tensor name: (lo, hi)}: per-dimension min and max of the offsets of the
    inputs that the tensor (store-normalised) depends on.

    This is the bounding box of ``get_overall_stencil_window(inputs, tensor)``
    (reference: src/soda/core.py:877-919) computed additively along the DAG.
    """
    bounds: Dict[str, Tuple[Tuple[int, ...], Tuple[int, ...]]] = {}
    zero = (0,) * self.dim
    for tensor in self.chronological_tensors:
      if tensor.is_input():
        bounds[tensor.name] = (zero, zero)
        continue
      lo = [None] * self.dim
      hi = [None] * self.dim
      for parent_name in tensor.parents:
        plo, phi = bounds[parent_name]
        for delta in tensor.ld_deltas(parent_name):
          for d in range(self.dim):
            a, b = plo[d] + delta[d], phi[d] + delta[d]
            lo[d] = a if lo[d] is None else min(lo[d], a)
            hi[d] = b if hi[d] is None else max(hi[d], b)
      if lo[0] is None:  # constant tensor
        lo, hi = list(zero), list(zero)
      bounds[tensor.name] = (tuple(lo), tuple(hi))
    return bounds

  def valid_box(self, name: str,
                extent: Sequence[int]) -> Tuple[Tuple[int, int], ...]:
    """Half-open index range per dimension where tensor ``name`` is defined on
    a grid of ``extent`` (reference: src/soda/codegen/frt/host.py:565-578)."""
    lo, hi = self.window_bounds[name]
    return tuple((max(0, -lo[d]), extent[d] - max(0, hi[d]))
                 for d in range(self.dim))

  def _calculate_stencil_window(self) -> None:
    window = get_overall_stencil_window(
        [self.tensors[name] for name in self.input_names],
        self.tensors[self.output_names[0]])
    distance = get_stencil_distance(window, self.tile_size)
    offset = distance - util.serialize(get_stencil_window_offset(window),
                                       self.tile_size)
    self.__dict__['_stencil_window'] = window
    self.__dict__['_stencil_distance'] = max(distance, offset)

  @property
  def stencil_distance(self) -> int:
    if '_stencil_distance' not in self.__dict__:
      self._calculate_stencil_window()
    return self.__dict__['_stencil_distance']

  @property
  def stencil_window(self):
    if '_stencil_window' not in self.__dict__:
      self._calculate_stencil_window()
    return self.__dict__['_stencil_window']

  @property
  def meta_lines(self) -> Tuple[str, ...]:
    lo, hi = self.window_bounds[self.output_names[0]]
    dims = tuple(h - l + 1 for l, h in zip(lo, hi))
    return (
        '// this file can be generated from the following SODA DSL',
        '/*\n%s\n*/' % self,
        '',
        '// stencil window size: %s' % (dims,),
        '',
    )

  # -- reuse analysis (becomes sliding-window sizes on the GPU) -----------------
  @_cached
  def reuse_buffers(self):
    """{producer name: [length, (start, end), ...]}
    (reference: src/soda/core.py:505-530,740-762)."""
    unroll_factor = self.unroll_factor
    self._reuse_buffer_lengths = {}
    reuse_buffers = {}
    for tensor in self.producer_tensors:
      reuse_buffer = _get_reuse_buffer(self.tile_size, tensor, unroll_factor)
      lengths = {}
      reuse_buffers[tensor.name] = reuse_buffer
      self._reuse_buffer_lengths[tensor.name] = lengths
      first = [True] * unroll_factor
      for start, end in reuse_buffer[1:]:
        if first[start % unroll_factor]:
          first[start % unroll_factor] = False
          if start >= unroll_factor:
            lengths[end] = end // unroll_factor
            continue
        lengths[end] = (end - start) // unroll_factor
    return reuse_buffers

  @property
  def reuse_buffer_lengths(self):
    self.reuse_buffers  # pylint: disable=pointless-statement
    return self._reuse_buffer_lengths

  @_cached
  def all_points(self):
    return {
        tensor.name: _get_points(self.tile_size, tensor, self.unroll_factor)
        for tensor in self.producer_tensors
    }

  @_cached
  def next_fifo(self):
    result = {}
    for name, reuse_buffer in self.reuse_buffers.items():
      result[name] = {
          start: end for start, end in reuse_buffer[1:] if start < end
      }
    return result


# ---------------------------------------------------------------------------
# free functions (same names as the reference's)
# ---------------------------------------------------------------------------


def _toposort(deps: Dict[str, set]) -> List[str]:
  """Level-by-level topological sort, alphabetical inside a level."""
  deps = {name: set(d) for name, d in deps.items()}
  order: List[str] = []
  while deps:
    ready = sorted(name for name, d in deps.items() if not d)
    if not ready:
      raise util.SemanticError('cyclic dependency among tensors: %s' %
                               ', '.join(sorted(deps)))
    order.extend(ready)
    for name in ready:
      del deps[name]
    done = set(ready)
    for d in deps.values():
      d -= done
  return order


def _unrolled_offsets(tensor, child, unroll_factor):
  offsets = set()
  for unroll_idx in range(unroll_factor):
    for offset in child.ld_offsets[tensor.name]:
      offsets.add(unroll_idx + child.st_offset - offset +
                  child.produce_offset - tensor.produce_offset)
  return offsets


def _get_reuse_chains(tile_size, tensor, unroll_factor):
  """One sorted tuple of stream offsets (distance behind the newest element)
  per unroll lane (reference: src/soda/core.py:684-725)."""
  accessed = set()
  for child in tensor.children.values():
    accessed |= _unrolled_offsets(tensor, child, unroll_factor)
  return [
      tuple(sorted(o for o in accessed if o % unroll_factor == lane))
      for lane in reversed(range(unroll_factor))
  ]


def _get_points(tile_size, tensor, unroll_factor):
  """{child: {offset: {unroll idx: index of the load among the child's
  loads}}} (reference: src/soda/core.py:728-762)."""
  all_points = {}
  for child in tensor.children.values():
    table = all_points[child.name] = {}
    for unroll_idx in range(unroll_factor):
      for idx, offset in enumerate(child.ld_offsets[tensor.name]):
        key = (unroll_idx + child.st_offset - offset + child.produce_offset -
               tensor.produce_offset)
        table.setdefault(key, {})[unroll_factor - 1 - unroll_idx] = idx
  return all_points


def _get_reuse_buffer(tile_size, tensor, unroll_factor):
  """[capacity, (start, end), ...] (reference: src/soda/core.py:765-795)."""
  reuse_buffer = [None]
  offsets = []
  for chain_id, chain in enumerate(
      _get_reuse_chains(tile_size, tensor, unroll_factor)):
    reuse_buffer.append((unroll_factor - 1 - chain_id, chain[0]))
    offsets.append(chain[0])
    for a, b in zip(chain, chain[1:]):
      reuse_buffer.append((a, b))
      offsets.append(b)
  reuse_buffer[0] = max(offsets) + 1
  return reuse_buffer


def get_indices_id(indices) -> str:
  return '_'.join(str(idx).replace('-', 'm') for idx in indices)


def get_stencil_distance(stencil_window, tile_size) -> int:
  return (max(util.serialize_iter(stencil_window, tile_size)) +
          util.serialize(get_stencil_window_offset(stencil_window), tile_size))


def get_stencil_dim(points) -> List[int]:
  points = list(points)
  dim = len(points[0])
  return [
      max(p[d] for p in points) - min(p[d] for p in points) + 1
      for d in range(dim)
  ]


def get_overall_stencil_window(input_tensor, output_tensor):
  """All offsets of ``input_tensor`` that ``output_tensor`` (store index
  normalised to 0) depends on; the union over inputs if an iterable is given
  (reference: src/soda/core.py:877-919)."""
  if isinstance(input_tensor, collections.abc.Iterable):
    points = set()
    for one_input in input_tensor:
      points |= set(get_overall_stencil_window(one_input, output_tensor))
    return tuple(sorted(points))

  cache: Dict[str, frozenset] = {}

  def window(tensor) -> frozenset:
    if tensor.name in cache:
      return cache[tensor.name]
    points = set()
    for name in tensor.parents:
      deltas = tensor.ld_deltas(name)
      if name == input_tensor.name:
        points.update(deltas)
      else:
        inner = window(tensor.parents[name])
        for delta in deltas:
          points.update(
              tuple(a + b for a, b in zip(point, delta)) for point in inner)
    cache[tensor.name] = frozenset(points)
    return cache[tensor.name]

  return tuple(sorted(window(output_tensor)))


def get_stencil_window_offset(stencil_window) -> Tuple[int, ...]:
  """Distance from the window's low corner to the store point."""
  points = list(stencil_window)
  return tuple(-min(p[d] for p in points) for d in range(len(points[0])))
