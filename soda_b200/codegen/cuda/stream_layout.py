"""The reference's stream data layout on the GPU: tiler / un-tiler.

The reference's generated host wrapper ``soda::app::<app>`` converts between
dense user arrays and the per-bank, burst-aligned, tile-by-tile stream buffers
the FPGA kernel consumes (reference: src/soda/codegen/frt/host.py:112-249 and
:340-427, docs/data-layout.md).  This module derives the layout constants from
a ``Stencil`` exactly as that wrapper does and runs the two transformations as
CUDA gather kernels through the C ABI of ``include/soda_layout.h``
(``soda_b200/csrc/soda_layout.cu``), on device memory.

There is no CPU path: the library is loaded with ctypes and every call fails
with ``SodaLayoutError`` when the library or a CUDA device is missing.
"""
import ctypes
from typing import List, Optional, Sequence

import numpy as np

from soda_b200 import core, util
from soda_b200.codegen.cuda import build as cuda_build

MAX_DIM = 3
MAX_BANKS = 32


class SodaLayoutError(RuntimeError):
  pass


class CLayout(ctypes.Structure):
  """``soda_stream_layout`` of include/soda_layout.h."""
  _fields_ = [
      ('struct_size', ctypes.c_int32),
      ('dim', ctypes.c_int32),
      ('elem_bytes', ctypes.c_int32),
      ('banks', ctypes.c_int32),
      ('extent', ctypes.c_int32 * MAX_DIM),
      ('stride', ctypes.c_int64 * MAX_DIM),
      ('tile_size', ctypes.c_int32 * MAX_DIM),
      ('stencil_dim', ctypes.c_int32 * MAX_DIM),
      ('window_offset', ctypes.c_int32 * MAX_DIM),
      ('window_dim', ctypes.c_int32 * MAX_DIM),
      ('stencil_distance', ctypes.c_int64),
      ('stencil_offset', ctypes.c_int64),
      ('produce_offset', ctypes.c_int64),
      ('elem_count_aligned_per_tile', ctypes.c_int64),
      ('elem_count_per_cycle', ctypes.c_int64),
  ]


def _round_up(a: int, b: int) -> int:
  return ((a - 1) // b + 1) * b


class TensorLayout:
  """Layout constants of one input or output tensor of ``stencil`` on a grid
  of ``extent`` cells, named like the generated C++.

  ``tile_size`` / ``burst_width`` default to the program's (the reference's
  wrapper takes them as run-time arguments with those defaults, reference:
  src/soda/codegen/frt/host.py:80-88).
  """

  def __init__(self, stencil, name: str, extent: Sequence[int],
               tile_size: Optional[Sequence[int]] = None,
               burst_width: Optional[int] = None,
               produce_offset: int = 0,
               stride: Optional[Sequence[int]] = None):
    dim = stencil.dim
    if dim not in (2, 3):
      raise util.SemanticError('stream layouts are implemented for 2-D and 3-D')
    if len(extent) != dim:
      raise util.SemanticError('extent must have %d entries' % dim)
    stmts = {s.name: s for s in stencil.input_stmts + stencil.output_stmts}
    if name not in stmts:
      raise util.SemanticError('%s is not an input or output tensor' % name)
    stmt = stmts[name]
    self.name = name
    self.is_output = name in stencil.output_names
    self.dim = dim
    self.extent = tuple(int(e) for e in extent)
    self.tile_size = tuple(tile_size or stencil.tile_size[:dim - 1])
    self.burst_width = int(burst_width or stencil.burst_width)
    self.elem_bits = stmt.haoda_type.width_in_bits
    self.banks = len(stmt.dram)
    self.dtype = np.dtype(stmt.haoda_type.numpy_name) if hasattr(
        stmt.haoda_type, 'numpy_name') else None
    if self.burst_width % self.elem_bits or self.elem_bits % 8:
      raise util.SemanticError('burst width must be a multiple of the element '
                               'width, elements whole bytes')
    # app constants (src/soda/codegen/frt/host.py:686-689)
    self.stencil_dim = tuple(core.get_stencil_dim(stencil.stencil_window))
    self.stencil_distance = stencil.stencil_distance
    first_in = stencil.input_stmts[0]
    first_out = stencil.output_stmts[0]
    tensors = stencil.tensors
    # un-tiler loop bounds: window first input -> first output (:352-356)
    window = core.get_overall_stencil_window(tensors[first_in.name],
                                             tensors[first_out.name])
    self.window_offset = tuple(core.get_stencil_window_offset(window))
    self.window_dim = tuple(core.get_stencil_dim(window))
    if self.is_output:
      # :396-403
      own = core.get_overall_stencil_window(
          [tensors[n] for n in stencil.input_names], tensors[name])
      distance = core.get_stencil_distance(own, stencil.tile_size)
      self.stencil_offset = distance - util.serialize(
          core.get_stencil_window_offset(own),
          tuple(self.tile_size) + (0,))
      self.produce_offset = 0
    else:
      self.stencil_offset = 0
      self.produce_offset = int(produce_offset)
    # run-time constants (:117-145)
    self.elem_count_per_cycle = self.burst_width // self.elem_bits * self.banks
    first_epc = self.burst_width // first_in.haoda_type.width_in_bits * len(
        first_in.dram)
    ref = first_out if self.is_output else first_in
    ref_epc = self.burst_width // ref.haoda_type.width_in_bits * len(ref.dram)
    per_tile = self.extent[-1]
    for size in self.tile_size:
      per_tile *= size
    cycles = (per_tile - 1) // first_epc + 1
    self.elem_count_aligned_per_tile = cycles * ref_epc
    if self.elem_count_aligned_per_tile < per_tile:
      # src/soda/codegen/frt/host.py:138-145 sizes the output tiles with the
      # first input's cycle count: fewer elements per cycle on the output side
      # would make the reference's tiles overlap in the buffer
      raise util.SemanticError(
          'outputs must move at least as many elements per cycle as the first '
          'input (%d < %d elements per tile)' %
          (self.elem_count_aligned_per_tile, per_tile))
    self.tile_count = [
        (self.extent[d] - self.stencil_dim[d] + 1 - 1) //
        (self.tile_size[d] - self.stencil_dim[d] + 1) + 1
        for d in range(dim - 1)
    ]
    for d in range(dim - 1):
      if self.tile_size[d] < self.stencil_dim[d] or \
          self.extent[d] < self.stencil_dim[d]:
        raise util.SemanticError(
            'tile size and extent must cover the stencil window in dimension '
            '%d' % d)
    tiles = 1
    for c in self.tile_count:
      tiles *= c
    self.elems_per_bank = (tiles * self.elem_count_aligned_per_tile + _round_up(
        self.stencil_distance, self.elem_count_per_cycle)) // self.banks
    if stride is None:
      stride = [1]
      for d in range(1, dim):
        stride.append(stride[-1] * self.extent[d - 1])
    self.stride = tuple(int(s) for s in stride)

  def c_struct(self) -> CLayout:
    c = CLayout()
    c.struct_size = ctypes.sizeof(CLayout)
    c.dim = self.dim
    c.elem_bytes = self.elem_bits // 8
    c.banks = self.banks
    for d in range(self.dim):
      c.extent[d] = self.extent[d]
      c.stride[d] = self.stride[d]
      c.stencil_dim[d] = self.stencil_dim[d]
      c.window_offset[d] = self.window_offset[d]
      c.window_dim[d] = self.window_dim[d]
    for d in range(self.dim - 1):
      c.tile_size[d] = self.tile_size[d]
    c.stencil_distance = self.stencil_distance
    c.stencil_offset = self.stencil_offset
    c.produce_offset = self.produce_offset
    c.elem_count_aligned_per_tile = self.elem_count_aligned_per_tile
    c.elem_count_per_cycle = self.elem_count_per_cycle
    return c


class LayoutLibrary:
  """ctypes binding of libsoda_layout (include/soda_layout.h)."""
  SYMBOLS = ('soda_layout_bank_elems', 'soda_layout_pack_device',
             'soda_layout_unpack_device', 'soda_layout_launch_count',
             'soda_layout_last_error')

  def __init__(self, path: Optional[str] = None):
    self.path = path or cuda_build.build_layout_library()
    try:
      self.lib = ctypes.CDLL(self.path)
    except OSError as e:
      raise SodaLayoutError('cannot load %s: %s' % (self.path, e))
    lib = self.lib
    lib.soda_layout_bank_elems.restype = ctypes.c_int
    lib.soda_layout_bank_elems.argtypes = [ctypes.POINTER(CLayout),
                                           ctypes.POINTER(ctypes.c_int64)]
    lib.soda_layout_pack_device.restype = ctypes.c_int
    lib.soda_layout_pack_device.argtypes = [
        ctypes.POINTER(CLayout), ctypes.c_void_p,
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]
    lib.soda_layout_unpack_device.restype = ctypes.c_int
    lib.soda_layout_unpack_device.argtypes = [
        ctypes.POINTER(CLayout), ctypes.POINTER(ctypes.c_void_p),
        ctypes.c_void_p, ctypes.c_void_p]
    lib.soda_layout_launch_count.restype = ctypes.c_int64
    lib.soda_layout_last_error.restype = ctypes.c_char_p

  def _check(self, status: int) -> None:
    if status != 0:
      message = self.lib.soda_layout_last_error() or b''
      raise SodaLayoutError('soda_layout status %d: %s' %
                            (status, message.decode()))

  def bank_elems(self, layout: TensorLayout) -> int:
    c = layout.c_struct()
    elems = ctypes.c_int64(0)
    self._check(self.lib.soda_layout_bank_elems(ctypes.byref(c),
                                                ctypes.byref(elems)))
    return int(elems.value)

  def launch_count(self) -> int:
    return int(self.lib.soda_layout_launch_count())

  def pack_device(self, layout: TensorLayout, dense_ptr: int,
                  bank_ptrs: Sequence[int], stream: int = 0) -> None:
    """Tiler on device pointers (asynchronous on ``stream``)."""
    if len(bank_ptrs) != layout.banks:
      raise SodaLayoutError('expected %d bank buffers' % layout.banks)
    c = layout.c_struct()
    banks = (ctypes.c_void_p * layout.banks)(*bank_ptrs)
    self._check(self.lib.soda_layout_pack_device(
        ctypes.byref(c), ctypes.c_void_p(dense_ptr), banks,
        ctypes.c_void_p(stream)))

  def unpack_device(self, layout: TensorLayout, bank_ptrs: Sequence[int],
                    dense_ptr: int, stream: int = 0) -> None:
    """Un-tiler on device pointers (asynchronous on ``stream``)."""
    if len(bank_ptrs) != layout.banks:
      raise SodaLayoutError('expected %d bank buffers' % layout.banks)
    c = layout.c_struct()
    banks = (ctypes.c_void_p * layout.banks)(*bank_ptrs)
    self._check(self.lib.soda_layout_unpack_device(
        ctypes.byref(c), banks, ctypes.c_void_p(dense_ptr),
        ctypes.c_void_p(stream)))

  # -- conveniences on torch tensors (device memory, streams: plumbing) ----------
  def pack(self, layout: TensorLayout, dense) -> List:
    """``dense``: a CUDA torch tensor of shape extent[::-1]; returns the bank
    buffers as 1-D CUDA tensors of the same dtype."""
    import torch
    if not dense.is_cuda or not dense.is_contiguous():
      raise SodaLayoutError('pack() needs a contiguous CUDA tensor')
    if dense.element_size() * 8 != layout.elem_bits or \
        tuple(dense.shape) != tuple(layout.extent[::-1]):
      raise SodaLayoutError('tensor does not match the layout')
    banks = [torch.empty(layout.elems_per_bank, dtype=dense.dtype,
                         device=dense.device) for _ in range(layout.banks)]
    self.pack_device(layout, dense.data_ptr(), [b.data_ptr() for b in banks],
                     torch.cuda.current_stream(dense.device).cuda_stream)
    return banks

  def unpack(self, layout: TensorLayout, banks: Sequence, dense) -> None:
    """Writes the valid interior of ``dense`` (CUDA tensor) from ``banks``."""
    import torch
    if not dense.is_cuda or not dense.is_contiguous():
      raise SodaLayoutError('unpack() needs a contiguous CUDA tensor')
    for bank in banks:
      if bank.numel() < layout.elems_per_bank or bank.dtype != dense.dtype:
        raise SodaLayoutError('bank buffer too small or of the wrong dtype')
    self.unpack_device(layout, [b.data_ptr() for b in banks], dense.data_ptr(),
                       torch.cuda.current_stream(dense.device).cuda_stream)
