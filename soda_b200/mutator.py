"""Index shifting / normalisation / CSE substitution over the IR.

Same entry points as the reference's ``soda.mutator``
(reference: src/soda/mutator.py:24-129).
"""
import collections
import operator
import types
from typing import Iterable, Mapping, MutableMapping, Optional, Tuple, Union

from soda_b200 import ir
from soda_b200 import visitor as soda_visitor


def shift(obj, offset, excluded=(), op=operator.sub):
  """Applies ``op(idx, offset)`` point-wise to every Ref not in ``excluded``.

  IR nodes are rebuilt (the argument is untouched); tensors are mutated in
  place and returned.
  """
  from soda_b200 import tensor

  def move(node, args):
    if isinstance(node, ir.Ref) and node.name not in excluded:
      node.idx = tuple(op(a, b) for a, b in zip(node.idx, offset))

  if isinstance(obj, ir.Node):
    return obj.visit(move)
  if isinstance(obj, tensor.Tensor):
    obj.mutate(move)
    return obj
  raise TypeError('argument is not an IR node or a tensor')


def normalize(obj: Union[ir.Node, Iterable[ir.Node]],
              references: Optional[Mapping[str, Tuple[int, ...]]] = None):
  """Shifts so that the least accessed index becomes 0."""
  if isinstance(obj, types.GeneratorType):
    return normalize(tuple(obj), references)
  norm_idx = soda_visitor.get_normalize_index(obj, references)
  if not any(norm_idx):
    return obj
  if isinstance(obj, ir.Node):
    return shift(obj, norm_idx)
  if isinstance(obj, collections.abc.Iterable):
    return type(obj)(shift(x, norm_idx) for x in obj)
  raise TypeError('argument is not an ir.Node or an iterable of ir.Nodes')


def replace_expressions(obj: ir.Node,
                        cses: MutableMapping[ir.Node, ir.Ref],
                        used: Optional[MutableMapping[ir.Node, ir.Node]] = None,
                        references: Optional[Mapping[str, Tuple[int,
                                                                 ...]]] = None):
  """Replaces every sub-tree whose normalised form is a key of ``cses`` by the
  mapped Ref, shifted back to where the sub-tree sat."""

  def substitute(node, args):
    norm_idx = soda_visitor.get_normalize_index(node, references)
    normalized = shift(node, norm_idx) if any(norm_idx) else node
    if normalized in cses:
      if used is not None and normalized not in used:
        used[normalized] = replace_expressions(
            normalized, {k: v for k, v in cses.items() if k != normalized},
            used)
      if any(norm_idx):
        return shift(cses[normalized], norm_idx, op=operator.add)
      return cses[normalized]
    return node

  return obj.visit(substitute)
