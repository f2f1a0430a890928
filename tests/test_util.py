"""Restates src/tests/test_util.py:7-11 of the reference and pins serialize
against the worked example in docs/data-layout.md:18-25."""
from soda_b200 import util


def test_deserialize_roundtrip():
  idx = (42, 23, 233)
  tile_size = (2333, 233, 0)
  assert idx == tuple(util.deserialize(util.serialize(idx, tile_size),
                                       tile_size))


def test_serialize_dim0_fastest():
  assert util.serialize((3, 0), (100, 0)) == 3
  assert util.serialize((3, 2), (100, 0)) == 203
  assert util.serialize((1, 2, 3), (10, 20, 0)) == 1 + 2 * 10 + 3 * 200
  assert util.serialize_iter([(0, 0), (1, 1)], (7, 0)) == [0, 8]
