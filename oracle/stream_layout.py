"""TEST INFRASTRUCTURE (oracle) - never imported by the product path.

CPU restatement of the reference's host-side *stream data layout*: how the
generated host wrapper ``soda::app::<app>`` tiles a dense user array into the
per-bank burst-aligned buffers the FPGA kernel streams, and how it un-tiles
the kernel's output buffers back (reference:
src/soda/codegen/frt/host.py:112-249 tiler, :340-427 un-tiler; the layout is
described with pictures in docs/data-layout.md, which the tests use as golden
vectors).

Pure-Python loops over (tile, coordinates in tile) exactly like the generated
C++ loops, so only use it at small sizes.

One deliberate deviation, flagged here because parity is judged on it: the
tiler computes the original coordinate as ``tile_index * (tile_size -
kStencilDim) + i`` (src/soda/codegen/frt/host.py:224-227) whereas the tile
count (:125-128), the last tile's size (:187-191), the un-tiler (:388-391) and
docs/data-layout.md:148-160 (second tile of a 150-wide image on a (100,*)
kernel with a 3x3 window starts at column 98 = 100 - 3 + 1) all use
``tile_size - kStencilDim + 1``.  With the tiler's literal formula every tile
after the first would be fed columns shifted by ``tile_index``; we follow the
documented layout (+1), which coincides with the literal code for single-tile
grids (the only case the reference's own tests run).

A quirk kept as is: the un-tiler takes its loop bounds from the window of the
*first* input (src/soda/codegen/frt/host.py:352-356).  For a program whose first
input has a smaller window than another one (denoise: ``f`` is read at one
point, ``u`` in a 5x5 window) every tile after the first therefore also writes
its leading columns, computed from cells outside the tile
(tests/test_stream_kernel.py::test_multi_input_seam_quirk).
"""
import dataclasses
from typing import List, Sequence, Tuple

import numpy as np


def round_up(a: int, b: int) -> int:
  """reference: src/soda/codegen/frt/host.py:115-116."""
  return ((a - 1) // b + 1) * b


@dataclasses.dataclass
class StreamLayout:
  """Run-time and app constants of one (tensor, grid) pair.

  extent           user array extent per dimension (dim 0 contiguous)
  tile_size        kernel tile size of dims 0 .. dim-2
  stencil_dim      kStencilDim<d>: size of the overall stencil window box
                   (src/soda/codegen/frt/host.py:686-688)
  stencil_distance kStencilDistance (:689)
  window_offset /  get_stencil_window_offset / get_stencil_dim of the window
  window_dim       first input -> first output (:352-356), per dimension
  stencil_offset   distance - serialize(window offset) of the output tensor
                   (:396-403); 0 for inputs
  elem_bits        element width in bits
  banks            number of DRAM banks of the tensor (len(stmt.dram))
  burst_width      bits per burst
  ref_elem_bits / ref_banks
                   the same two of the *first input* (tiles are aligned with its
                   elem_count_per_cycle for inputs, :138-143) or of the *first
                   output* (for outputs, :144-145)
  produce_offset   input only: stream offset at which the tensor is produced
                   relative to the first input (:245); 0 for single-input
                   programs
  """
  extent: Tuple[int, ...]
  tile_size: Tuple[int, ...]
  stencil_dim: Tuple[int, ...]
  stencil_distance: int
  window_offset: Tuple[int, ...]
  window_dim: Tuple[int, ...]
  stencil_offset: int
  elem_bits: int
  banks: int
  burst_width: int
  ref_elem_bits: int
  ref_banks: int
  first_input_elem_bits: int
  first_input_banks: int
  produce_offset: int = 0

  @property
  def dim(self) -> int:
    return len(self.extent)

  @property
  def elem_count_per_cycle(self) -> int:  # of this tensor (:117-120)
    return self.burst_width // self.elem_bits * self.banks

  @property
  def ref_elem_count_per_cycle(self) -> int:
    return self.burst_width // self.ref_elem_bits * self.ref_banks

  @property
  def tile_count(self) -> List[int]:  # :124-128
    return [(self.extent[d] - self.stencil_dim[d] + 1 - 1) //
            (self.tile_size[d] - self.stencil_dim[d] + 1) + 1
            for d in range(self.dim - 1)]

  @property
  def elem_count_aligned_per_tile(self) -> int:  # :134-145
    per_tile = self.extent[-1]
    for d in range(self.dim - 1):
      per_tile *= self.tile_size[d]
    # cycle_count_per_tile always uses the first input's elements per cycle
    first = self.burst_width // self.first_input_elem_bits * \
        self.first_input_banks
    cycles = (per_tile - 1) // first + 1
    return cycles * self.ref_elem_count_per_cycle

  @property
  def elems_per_bank(self) -> int:  # buf_size / sizeof, :147-162
    tiles = 1
    for c in self.tile_count:
      tiles *= c
    return (tiles * self.elem_count_aligned_per_tile + round_up(
        self.stencil_distance, self.elem_count_per_cycle)) // self.banks

  def tile_stride(self, d: int) -> int:
    return self.tile_size[d] - self.stencil_dim[d] + 1

  def actual_tile_size(self, d: int, index: int) -> int:  # :187-191
    if index == self.tile_count[d] - 1:
      return self.extent[d] - self.tile_stride(d) * index
    return self.tile_size[d]


def _tile_indices(layout: StreamLayout):
  """Yields (tile index per dim, linear tile index), dim 0 fastest (:235-238)."""
  counts = layout.tile_count
  total = 1
  for c in counts:
    total *= c
  for linear in range(total):
    rest = linear
    index = []
    for c in counts:
      index.append(rest % c)
      rest //= c
    yield index, linear


def _offset_in_tile(coords: Sequence[int], tile_size: Sequence[int]) -> int:
  offset, pitch = 0, 1
  for d, c in enumerate(coords):
    offset += c * pitch
    if d < len(tile_size):
      pitch *= tile_size[d]
  return offset


def tile(layout: StreamLayout, dense: np.ndarray, void=0) -> List[np.ndarray]:
  """Dense array (numpy shape = extent reversed) -> one 1-D buffer per bank.
  Positions the reference leaves uninitialised are set to ``void``.
  reference: src/soda/codegen/frt/host.py:181-249."""
  flat = dense.reshape(-1)
  banks = [np.full(layout.elems_per_bank, void, dtype=dense.dtype)
           for _ in range(layout.banks)]
  dim = layout.dim
  strides = [1]
  for d in range(1, dim):
    strides.append(strides[-1] * layout.extent[d - 1])
  aligned = layout.elem_count_aligned_per_tile
  for index, linear in _tile_indices(layout):
    sizes = [layout.actual_tile_size(d, index[d]) for d in range(dim - 1)]
    ranges = [range(s) for s in sizes] + [range(layout.extent[-1])]
    for coords in np.ndindex(*[len(r) for r in reversed(ranges)]):
      coords = coords[::-1]  # dim 0 first
      tiled = linear * aligned + _offset_in_tile(coords, layout.tile_size)
      original = 0
      for d in range(dim - 1):
        original += (index[d] * layout.tile_stride(d) + coords[d]) * strides[d]
      original += coords[-1] * strides[-1]
      banks[tiled % layout.banks][tiled // layout.banks] = \
          flat[max(0, original - layout.produce_offset)]
  return banks


def untile(layout: StreamLayout, banks: Sequence[np.ndarray],
           dense: np.ndarray) -> np.ndarray:
  """Per-bank output buffers -> the valid interior of ``dense`` (everything
  else keeps its old contents).
  reference: src/soda/codegen/frt/host.py:340-427."""
  flat = dense.reshape(-1)
  dim = layout.dim
  strides = [1]
  for d in range(1, dim):
    strides.append(strides[-1] * layout.extent[d - 1])
  aligned = layout.elem_count_aligned_per_tile
  lo = [max(0, layout.window_offset[d]) for d in range(dim)]
  cut = [max(0, layout.window_dim[d] - 1 - layout.window_offset[d])
         for d in range(dim)]
  for index, linear in _tile_indices(layout):
    sizes = [layout.actual_tile_size(d, index[d]) for d in range(dim - 1)]
    ranges = [range(lo[d], sizes[d] - cut[d]) for d in range(dim - 1)]
    ranges.append(range(lo[-1], layout.extent[-1] - cut[-1]))
    if any(len(r) == 0 for r in ranges):
      continue
    for pick in np.ndindex(*[len(r) for r in reversed(ranges)]):
      pick = pick[::-1]
      coords = [ranges[d][pick[d]] for d in range(dim)]
      tiled = linear * aligned + _offset_in_tile(coords, layout.tile_size) + \
          layout.stencil_offset
      original = 0
      for d in range(dim - 1):
        original += (index[d] * layout.tile_stride(d) + coords[d]) * strides[d]
      original += coords[-1] * strides[-1]
      flat[original] = banks[tiled % layout.banks][tiled // layout.banks]
  return dense
