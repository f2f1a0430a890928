"""Index helpers and error types shared by the SODA front end.

Mirrors the public surface of the reference's ``soda.util``
(reference: src/soda/util.py:4-24) plus the handful of ``haoda.util`` names the
front end relies on (SemanticError, InternalError, InputError, idx2str,
lst2str).  ``haoda`` is not vendored in the reference tree, so these are
re-stated here from their call sites.
"""
import functools
import operator
from typing import Iterable, Iterator, Sequence, Tuple

COORDS_TILED = 'xyzw'
COORDS_IN_TILE = 'ijkl'
COORDS_IN_ORIG = 'pqrs'
MAX_DRAM_BANK = 4


class SemanticError(Exception):
  """The program is syntactically fine but has no meaning (e.g. iterate: 0)."""


class SemanticWarn(Exception):
  pass


class InternalError(Exception):
  pass


class InputError(Exception):
  pass


def serialize(vec: Sequence[int], tile_size: Sequence[int]) -> int:
  """Row-major-with-dim-0-fastest linear offset of ``vec`` inside a tile.

  The last entry of ``tile_size`` is never used (the last dimension is the
  unbounded streaming dimension).  reference: src/soda/util.py:9-12.
  """
  offset = vec[0]
  pitch = 1
  for dim in range(1, len(tile_size)):
    pitch *= tile_size[dim - 1]
    offset += vec[dim] * pitch
  return offset


def serialize_iter(iterative: Iterable[Sequence[int]],
                   tile_size: Sequence[int]):
  return [serialize(x, tile_size) for x in iterative]


def deserialize_generator(offset: int,
                          tile_size: Sequence[int]) -> Iterator[int]:
  for size in tile_size[:-1]:
    yield offset % size
    offset = offset // size
  yield offset


def deserialize(offset: int, tile_size: Sequence[int]) -> Tuple[int, ...]:
  """Inverse of :func:`serialize` (reference: src/soda/util.py:17-24)."""
  return tuple(deserialize_generator(offset, tile_size))


def idx2str(idx) -> str:
  return '(%s)' % ', '.join(map(str, idx))


def lst2str(lst) -> str:
  return '[%s]' % ', '.join(map(str, lst))


def product(values: Iterable[int]) -> int:
  return functools.reduce(operator.mul, values, 1)
