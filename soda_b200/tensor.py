"""Tensor: a node of the high-level stencil DAG (input, local, or output).

Mirrors the attribute surface of the reference's ``soda.tensor.Tensor``
(reference: src/soda/tensor.py:14-147): ``st_ref/st_idx/st_offset``,
``ld_refs/ld_indices/ld_offsets``, ``parents/children``, ``expr/lets``,
``haoda_type``; plus ``produce_offset/consume_offset/max_access`` filled in by
``core.Stencil.tensors``.
"""
import collections
import copy
from typing import Dict, Tuple

from soda_b200 import grammar, ir, util


class Tensor:

  def __init__(self, stmt, tile_size):
    self.haoda_type = stmt.haoda_type
    self._tile_size = tuple(tile_size)
    if isinstance(stmt, grammar.LocalStmtOrOutputStmt):
      self.st_ref = copy.copy(stmt.ref)
      self.st_ref.parent = self
      self.lets = tuple(stmt.let)
      self.expr = stmt.expr
      self._name = None
    elif isinstance(stmt, grammar.InputStmt):
      self._name = stmt.name
      self.st_ref = None
      self.lets = ()
      self.expr = None
    else:
      raise util.InternalError('cannot initialize a Tensor from %s' %
                               type(stmt))
    # wired up by Stencil.tensors
    self.parents: Dict[str, 'Tensor'] = collections.OrderedDict()
    self.children: Dict[str, 'Tensor'] = collections.OrderedDict()
    self.ld_refs: Dict[str, list] = collections.OrderedDict()
    self.produce_offset = 0
    self.consume_offset = 0
    self.max_access = 0

  @property
  def name(self) -> str:
    return self.st_ref.name if self.st_ref is not None else self._name

  @property
  def st_idx(self) -> Tuple[int, ...]:
    if self.st_ref is not None:
      return self.st_ref.idx
    return (0,) * len(self._tile_size)

  @property
  def st_offset(self) -> int:
    return util.serialize(self.st_idx, self._tile_size)

  @property
  def ld_indices(self):
    """{parent name: {accessed index tuple: Ref}} in serialised-offset order."""
    return collections.OrderedDict(
        (name, collections.OrderedDict((ref.idx, ref)
                                       for ref in refs))
        for name, refs in self.ld_refs.items())

  @property
  def ld_offsets(self):
    """{parent name: {serialised offset: Ref}}."""
    return collections.OrderedDict(
        (name,
         collections.OrderedDict(
             (util.serialize(ref.idx, self._tile_size), ref) for ref in refs))
        for name, refs in self.ld_refs.items())

  def ld_deltas(self, parent_name: str):
    """Relative accesses ``ld_idx - st_idx`` into ``parent_name``
    (the address rule of the reference's golden loops, reference:
    src/soda/codegen/frt/host.py:587-592)."""
    return tuple(
        tuple(a - b
              for a, b in zip(ref.idx, self.st_idx))
        for ref in self.ld_refs[parent_name])

  @property
  def c_type(self) -> str:
    return self.haoda_type.c_type

  def mutate(self, callback, args=None) -> None:
    self.lets = tuple(let.visit(callback, args) for let in self.lets)
    self.expr = self.expr.visit(callback, args)
    self.st_ref = self.st_ref.visit(callback, args)

  def visit_loads(self, callback, args=None) -> None:
    for let in self.lets:
      let.visit(callback, args)
    if self.expr is not None:
      self.expr.visit(callback, args)

  def is_output(self) -> bool:
    return len(self.children) == 0

  def is_input(self) -> bool:
    return len(self.parents) == 0

  def is_producer(self) -> bool:
    return not self.is_output()

  def is_consumer(self) -> bool:
    return not self.is_input()

  def __str__(self) -> str:
    return ('Tensor\n  {}: {} = {}\n  store: {}\n  parents: {}\n'
            '  children: {}').format(self.haoda_type, self.name, self.expr,
                                     self.st_ref, util.idx2str(self.parents),
                                     util.idx2str(self.children))
