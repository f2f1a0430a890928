"""Load collectors over IR nodes and tensors.

Same entry points as the reference's ``soda.visitor``
(reference: src/soda/visitor.py:15-122).
"""
import collections
from typing import Dict, Iterable, List, Mapping, Optional, Tuple, Union

from soda_b200 import ir


def _walk(obj, callback, acc) -> None:
  # late import: tensor imports grammar which imports ir
  from soda_b200 import tensor
  if isinstance(obj, ir.Node):
    obj.visit(callback, acc)
  elif isinstance(obj, tensor.Tensor):
    obj.visit_loads(callback, acc)
  else:
    raise TypeError('argument is not an IR node or a tensor.Tensor')


def get_load_tuple(obj) -> Tuple[ir.Ref, ...]:
  """All load references, in source order, duplicates kept."""
  loads: List[ir.Ref] = []

  def collect(node, acc):
    if isinstance(node, ir.Ref):
      acc.append(node)

  _walk(obj, collect, loads)
  return tuple(loads)


def get_load_set(obj) -> Tuple[ir.Ref, ...]:
  """Unique load references, first-seen order."""
  loads: Dict[ir.Ref, None] = collections.OrderedDict()

  def collect(node, acc):
    if isinstance(node, ir.Ref):
      acc[node] = None

  _walk(obj, collect, loads)
  return tuple(loads)


def get_load_dict(obj) -> Dict[str, List[ir.Ref]]:
  """{tensor name: [loads of that tensor]} in source order."""
  loads: Dict[str, List[ir.Ref]] = collections.OrderedDict()

  def collect(node, acc):
    if isinstance(node, ir.Ref):
      acc.setdefault(node.name, []).append(node)

  _walk(obj, collect, loads)
  return loads


def get_normalize_index(
    obj: Union[ir.Node, Iterable[ir.Node]],
    references: Optional[Mapping[str, Tuple[int, ...]]] = None
) -> Tuple[int, ...]:
  """The index that, subtracted from every load, makes the least access 0.

  "Least" compares indices from the last dimension down (stream order).
  """
  if isinstance(obj, ir.Node):
    obj = (obj,)
  elif not isinstance(obj, collections.abc.Iterable):
    raise TypeError('argument is not an ir.Node or an iterable of ir.Nodes')

  def rel_idx(load: ir.Ref) -> Tuple[int, ...]:
    base = None if references is None else references.get(load.name)
    if base is None:
      return load.idx
    return tuple(x - y for x, y in zip(load.idx, base))

  loads = [load for node in obj for load in get_load_tuple(node)]
  if not loads:
    return ()
  return min((rel_idx(load) for load in loads),
             key=lambda idx: tuple(reversed(idx)))
