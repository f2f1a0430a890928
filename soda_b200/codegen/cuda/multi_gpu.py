"""Multi-GPU slab runtime: a thin ctypes layer over the slab entry points of
the C ABI (include/soda_cuda.h, ``soda_cuda_slab_*``; implementation:
csrc/soda_slab.cuh).

The grid is split along the outermost (streamed, ``*``) dimension into
contiguous slabs, one per rank - the dimension the reference itself treats as
unbounded (reference: README.md:223, src/soda/codegen/frt/host.py:124-131 tiles
every other dimension).  Groups of passes (one pass = ``time_block`` fused
iterations) run between two halo exchanges; the last pass of a group is issued
as three launches so that the exchange of the boundary slices overlaps the
interior.  Results are bit-identical to one GPU: every stored cell sees the
same dependency cone in the same operation order.

Everything that decides *what* runs (slab bounds, ghost depths, exchange
groups, store boxes, the chunked host pipeline) lives behind the ABI.  This
module only
  * wraps the library's device arrays as ``torch`` tensors so that callers can
    fill and read them,
  * supplies a transport: the library's own NCCL communicator (``'nccl'``; the
    unique id travels through ``torch.distributed``), or a callback that moves
    the halos with ``torch.distributed`` point-to-point operations
    (``'torch'``: gloo in the CPU tests, NCCL on GPUs).

The reference has no distributed path at all; this is
new functionality required by BASELINE.json's north star.
"""
import ctypes
import traceback
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from soda_b200.codegen.cuda import launcher

NCCL_ID_BYTES = 128


class HaloOp(ctypes.Structure):
  _fields_ = [
      ('send', ctypes.c_int32),
      ('peer', ctypes.c_int32),
      ('ptr', ctypes.c_void_p),
      ('bytes', ctypes.c_int64),
  ]


EXCHANGE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int32,
                               ctypes.POINTER(HaloOp), ctypes.c_int32,
                               ctypes.c_void_p)


class SlabOpts(ctypes.Structure):
  _fields_ = [
      ('struct_size', ctypes.c_int32),
      ('rank', ctypes.c_int32),
      ('world', ctypes.c_int32),
      ('device', ctypes.c_int32),
      ('stream', ctypes.c_void_p),
      ('segment', ctypes.c_int32),
      ('exchange_every', ctypes.c_int32),
      ('no_overlap', ctypes.c_int32),
      ('host_chunks', ctypes.c_int32),
      ('exchange', EXCHANGE_FN),
      ('exchange_user', ctypes.c_void_p),
      ('nccl_id', ctypes.c_void_p),
      ('reserved', ctypes.c_int32 * 8),
  ]


class SlabInfo(ctypes.Structure):
  _fields_ = [
      ('begin', ctypes.c_int32),
      ('end', ctypes.c_int32),
      ('local_begin', ctypes.c_int32),
      ('local_end', ctypes.c_int32),
      ('ghost_lo', ctypes.c_int32),
      ('ghost_hi', ctypes.c_int32),
      ('local_extent', ctypes.c_int32 * launcher.MAX_DIM),
      ('pitch', ctypes.c_int64 * 2),
      ('num_groups', ctypes.c_int32),
      ('group_passes', ctypes.c_int32 * 16),
  ]


def split_slices(total: int, world: int) -> List[Tuple[int, int]]:
  """Contiguous [begin, end) ranges of the streamed dimension per rank (the
  rule of csrc/soda_slab.cuh, ``split_slices``)."""
  base, rest = divmod(total, world)
  ranges = []
  begin = 0
  for rank in range(world):
    size = base + (1 if rank < rest else 0)
    ranges.append((begin, begin + size))
    begin += size
  return ranges


class _DeviceBytes:
  """Raw device memory as a ``__cuda_array_interface__`` provider."""

  def __init__(self, ptr: int, nbytes: int):
    self.__cuda_array_interface__ = {
        'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False),
        'version': 2, 'strides': None,
    }


def _bytes_tensor(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
  """``nbytes`` bytes at ``ptr`` as a uint8 tensor that aliases the memory."""
  if device.type == 'cuda':
    return torch.as_tensor(_DeviceBytes(ptr, nbytes), device=device)
  buffer = (ctypes.c_char * nbytes).from_address(ptr)
  return torch.frombuffer(buffer, dtype=torch.uint8)


class SlabRunner:
  """This rank's slab of a larger global grid.

  ``global_extent``: extent of the whole grid (dimension 0 first).
  ``self.inputs`` / ``self.outputs`` are torch tensors of shape
  ``(local_slices[, extent[1]], pitch)`` that alias the library's device
  arrays; ``self.own`` is the local slice range this rank owns (fill those of
  the inputs; ghost slices are refreshed by ``run``).

  ``transport``: ``'nccl'`` - the library's own NCCL communicator; ``'torch'`` -
  halos move through ``torch.distributed`` point-to-point operations of
  ``group``; ``'auto'`` - ``'nccl'`` on CUDA devices, ``'torch'`` otherwise.
  ``exchange_every``: passes per halo exchange (``None``: as many as keep the
  ghost below 2.5 % of the slab; ``-1``: all of them, which is what
  ``run_host`` needs).
  """

  def __init__(self,
               program: launcher.CudaProgram,
               global_extent: Sequence[int],
               device: torch.device,
               rank: Optional[int] = None,
               world: Optional[int] = None,
               group=None,
               stream_handle: int = 0,
               exchange_every: Optional[int] = None,
               transport: str = 'auto',
               overlap: bool = True,
               host_chunks: int = 0,
               edge_chunks: Optional[str] = None,
               segment: int = 0,
               dry_run: bool = False):
    self.program = program
    self.group = group
    self.rank = dist.get_rank(group) if rank is None else rank
    self.world = dist.get_world_size(group) if world is None else world
    self.device = torch.device(device)
    self.global_extent = tuple(global_extent)
    self.dim = program.dim
    self.stream_handle = stream_handle
    lib = program.lib
    for name in ('soda_cuda_slab_create', 'soda_cuda_slab_run'):
      if not hasattr(lib, name):
        raise launcher.SodaCudaError(
            -1, '%s does not export %s' % (program.lib_path, name))
    if transport == 'auto':
      transport = 'nccl' if self.device.type == 'cuda' else 'torch'
    self.transport = transport
    self._pending = []
    opts = SlabOpts()
    opts.struct_size = ctypes.sizeof(SlabOpts)
    opts.rank, opts.world = self.rank, self.world
    opts.device = self.device.index if self.device.type == 'cuda' and \
        self.device.index is not None else -1
    opts.stream = stream_handle or None
    opts.segment = segment
    opts.exchange_every = exchange_every or 0
    opts.no_overlap = 0 if overlap else 1
    opts.host_chunks = host_chunks
    opts.reserved[0] = 1 if dry_run else 0
    opts.reserved[1] = {None: 0, 'natural': 1, 'last': 2}[edge_chunks]
    self._callback = EXCHANGE_FN(self._exchange_callback)
    self._nccl_id = None
    if dry_run:
      pass
    elif self.world > 1 and transport == 'torch':
      opts.exchange = self._callback
    elif self.world > 1:
      self._nccl_id = self._shared_nccl_id()
      opts.nccl_id = ctypes.cast(self._nccl_id, ctypes.c_void_p)
    self._opts = opts
    self.handle = ctypes.c_void_p()
    c_extent = (ctypes.c_int32 * self.dim)(*self.global_extent)
    program._check(lib.soda_cuda_slab_create(c_extent, ctypes.byref(opts),
                                             ctypes.byref(self.handle)))
    info = SlabInfo()
    program._check(lib.soda_cuda_slab_get_info(self.handle, ctypes.byref(info)))
    self.info = info
    self.begin, self.end = info.begin, info.end
    self.local_begin, self.local_end = info.local_begin, info.local_end
    self.reach_lo, self.reach_hi = info.ghost_lo, info.ghost_hi
    self.local_extent = tuple(info.local_extent[d] for d in range(self.dim))
    self.own = (self.begin - self.local_begin, self.end - self.local_begin)
    self.pitch = int(info.pitch[0])
    s_dim = self.dim - 1
    self.pass_reach = []
    for index in range(program.num_passes):
      pinfo = program.pass_info(index)
      self.pass_reach.append((max(0, -pinfo.reach_lo[s_dim]),
                              max(0, pinfo.reach_hi[s_dim])))
    # groups as lists of pass indices (sizes from the library)
    sizes = [info.group_passes[g] for g in range(min(info.num_groups, 16))]
    self.num_groups = info.num_groups
    self.groups, first = [], 0
    for size in sizes:
      self.groups.append(list(range(first, first + size)))
      first += size
    self.inputs, self.outputs = [], []
    self._launches_at_start = program.launch_count()
    if dry_run:
      return
    n_in, n_out = len(program.input_dtypes), len(program.output_dtypes)
    d_in = (ctypes.c_void_p * max(1, n_in))()
    d_out = (ctypes.c_void_p * max(1, n_out))()
    program._check(lib.soda_cuda_slab_buffers(self.handle, d_in, d_out))
    self.inputs = [self._wrap(d_in[i], dt)
                   for i, dt in enumerate(program.input_dtypes)]
    self.outputs = [self._wrap(d_out[o], dt)
                    for o, dt in enumerate(program.output_dtypes)]

  # -- buffers -------------------------------------------------------------------
  def _shape(self):
    return tuple(self.local_extent[1:][::-1]) + (self.pitch,)

  def _wrap(self, ptr: int, np_dtype) -> torch.Tensor:
    shape = self._shape()
    nbytes = int(np.prod(shape)) * np.dtype(np_dtype).itemsize
    raw = _bytes_tensor(ptr, nbytes, self.device)
    return raw.view(getattr(torch, str(np_dtype))).view(shape)

  def view(self, tensor: torch.Tensor) -> torch.Tensor:
    """The un-padded part of a local array."""
    return tensor[..., :self.global_extent[0]]

  @property
  def launches(self) -> int:
    return self.program.launch_count() - self._launches_at_start

  # -- transports -----------------------------------------------------------------
  def _shared_nccl_id(self):
    """Rank 0 asks the library for an NCCL unique id; every rank gets it
    through ``torch.distributed``."""
    ident = (ctypes.c_char * NCCL_ID_BYTES)()
    payload = [None]
    if self.rank == 0:
      self.program._check(self.program.lib.soda_cuda_nccl_unique_id(ident))
      payload = [bytes(ident.raw)]
    src = 0 if self.group is None else dist.get_global_rank(self.group, 0)
    dist.broadcast_object_list(payload, src=src, group=self.group)
    ident.raw = payload[0]
    return ident

  def _peer(self, rank: int) -> int:
    if self.group is None:
      return rank
    return dist.get_global_rank(self.group, rank)

  def _stream_context(self, stream):
    if self.device.type == 'cuda' and stream:
      return torch.cuda.stream(torch.cuda.ExternalStream(stream,
                                                         device=self.device))
    import contextlib
    return contextlib.nullcontext()

  def _exchange_callback(self, user, phase, ops, num_ops, stream) -> int:
    """soda_cuda_exchange_fn over torch.distributed: phase 0 starts the
    transfers after what is queued on ``stream``, phase 1 makes ``stream``
    wait for them."""
    try:
      with self._stream_context(stream):
        if phase == 0:
          p2p = []
          for index in range(num_ops):
            op = ops[index]
            tensor = _bytes_tensor(op.ptr, op.bytes, self.device)
            p2p.append(dist.P2POp(dist.isend if op.send else dist.irecv, tensor,
                                  self._peer(op.peer), self.group))
          self._pending = dist.batch_isend_irecv(p2p) if p2p else []
        else:
          for work in self._pending:
            work.wait()
          self._pending = []
      return 0
    except Exception:  # pylint: disable=broad-except
      traceback.print_exc()
      return 1

  # -- running --------------------------------------------------------------------
  def run(self) -> None:
    """All ``iterate`` iterations: ``self.inputs`` -> ``self.outputs``; the
    ghost slices of the inputs are refreshed first, so callers only fill the
    slices they own.  Queued on the slab's stream."""
    self.program._check(self.program.lib.soda_cuda_slab_run(self.handle))

  def exchange(self, tensors=None, reach_lo: int = 0, reach_hi: int = 0) -> None:
    """One halo exchange of the input arrays (benchmarks time it alone)."""
    del tensors  # the library exchanges its own input arrays
    self.program._check(self.program.lib.soda_cuda_slab_exchange_inputs(
        self.handle, reach_lo, reach_hi))

  def run_host(self, inputs: Dict[str, np.ndarray],
               outputs: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Host arrays holding this rank's own slices in, the same slices of the
    outputs out, through the chunked H2D / passes / D2H pipeline (needs
    ``exchange_every=-1``)."""
    prog = self.program
    ins = [inputs[name] for name in prog.input_names]
    outs = [outputs[name] for name in prog.output_names]
    own_extent = self.global_extent[:-1] + (self.end - self.begin,)
    for array, dtype in zip(ins + outs, prog.input_dtypes + prog.output_dtypes):
      if array.dtype != dtype or tuple(array.shape[::-1]) != own_extent:
        raise TypeError('tensor must be %s of extent %s' % (dtype, own_extent))
    stride_ptr = ctypes.POINTER(ctypes.c_int32)
    strides = [(ctypes.c_int32 * prog.dim)(*prog._strides_of(a))
               for a in ins + outs]
    n_in, n_out = len(ins), len(outs)
    in_ptrs = (ctypes.c_void_p * n_in)(*[a.ctypes.data for a in ins])
    out_ptrs = (ctypes.c_void_p * n_out)(*[a.ctypes.data for a in outs])
    in_strides = (stride_ptr * n_in)(
        *[ctypes.cast(s, stride_ptr) for s in strides[:n_in]])
    out_strides = (stride_ptr * n_out)(
        *[ctypes.cast(s, stride_ptr) for s in strides[n_in:]])
    prog._check(prog.lib.soda_cuda_slab_run_host(self.handle, in_ptrs,
                                                 in_strides, out_ptrs,
                                                 out_strides))
    return outputs

  def close(self) -> None:
    if getattr(self, 'handle', None):
      self.inputs = self.outputs = []
      self.program.lib.soda_cuda_slab_destroy(self.handle)
      self.handle = ctypes.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:  # pylint: disable=broad-except
      pass


def run_host_multi(program: launcher.CudaProgram,
                   inputs: Dict[str, np.ndarray],
                   outputs: Optional[Dict[str, np.ndarray]] = None,
                   num_devices: int = 1,
                   devices: Optional[Sequence[int]] = None,
                   opts: Optional[launcher.Opts] = None
                   ) -> Dict[str, np.ndarray]:
  """One process, ``num_devices`` GPUs: the whole grid in host arrays
  (``soda_cuda_multi_run_host``; what ``sodac --cuda-gpus N`` compiles into the
  program-named entry point)."""
  prog = program
  ins = [inputs[name] for name in prog.input_names]
  extent = prog._extent_of(ins[0])
  if outputs is None:
    outputs = {name: np.zeros(extent[::-1], dtype=dtype)
               for name, dtype in zip(prog.output_names, prog.output_dtypes)}
  outs = [outputs[name] for name in prog.output_names]
  stride_ptr = ctypes.POINTER(ctypes.c_int32)
  strides = [(ctypes.c_int32 * prog.dim)(*prog._strides_of(a))
             for a in ins + outs]
  n_in, n_out = len(ins), len(outs)
  in_ptrs = (ctypes.c_void_p * n_in)(*[a.ctypes.data for a in ins])
  out_ptrs = (ctypes.c_void_p * n_out)(*[a.ctypes.data for a in outs])
  in_strides = (stride_ptr * n_in)(
      *[ctypes.cast(s, stride_ptr) for s in strides[:n_in]])
  out_strides = (stride_ptr * n_out)(
      *[ctypes.cast(s, stride_ptr) for s in strides[n_in:]])
  c_extent = (ctypes.c_int32 * prog.dim)(*extent)
  c_devices = None
  if devices is not None:
    c_devices = (ctypes.c_int32 * len(devices))(*devices)
    num_devices = len(devices)
  prog._check(prog.lib.soda_cuda_multi_run_host(
      in_ptrs, in_strides, out_ptrs, out_strides, c_extent, c_devices,
      num_devices, ctypes.byref(opts) if opts is not None else None))
  return outputs
