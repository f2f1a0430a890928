"""Multi-GPU slab runtime: one process per GPU, halo exchange per group of
passes.

The grid is split along the outermost (streamed, ``*``) dimension into
contiguous slabs, one per rank - the dimension the reference itself treats as
unbounded (reference: README.md:223, src/soda/codegen/frt/host.py:124-131 tiles
every other dimension).  Between passes (one pass = ``time_block`` fused
iterations) neighbouring ranks swap the ``reach`` slices next to their common
boundary with ``torch.distributed`` point-to-point operations (NCCL send/recv
over NVLink on GPUs, gloo in the CPU tests).  There is no other collective:
a stencil has no reduction.

Every rank's local array covers its own slices plus ``reach`` ghost slices on
each side that exist in the global grid; at the global border there is no
ghost, so the kernels' TMA loads zero-fill there exactly as on one GPU.  Every
stored value therefore has the same dependency cone and the same operation
order as in the single-GPU run: results are bit-identical.

Exchange groups: temporal blocking one level up.  A rank that holds
``k x reach`` ghost slices can run ``k`` passes without talking to anybody -
pass ``j`` of the group also computes the ghost slices the remaining passes
still need, exactly like a strip of the 2-D kernel recomputes its halo - and
then swaps ``k x reach`` slices at once.  The redundant work is
``k x reach / slab`` (0.4 % for the bench: 66 ghost rows per side of a
16384-row slab, one exchange per 64 iterations instead of eleven), the
synchronisations with the neighbours drop by ``k``.  ``exchange_every`` picks
``k`` (default: as many passes as keep the ghost below 2.5 % of the slab).

The reference has no distributed path at all; this file
is new functionality required by BASELINE.json's north star.
"""
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from soda_b200.codegen.cuda import launcher


def split_slices(total: int, world: int) -> List[Tuple[int, int]]:
  """Contiguous [begin, end) ranges of the streamed dimension per rank."""
  base, rest = divmod(total, world)
  ranges = []
  begin = 0
  for rank in range(world):
    size = base + (1 if rank < rest else 0)
    ranges.append((begin, begin + size))
    begin += size
  return ranges


class SlabRunner:
  """Runs a compiled program on this rank's slab of a larger global grid.

  ``global_extent``: extent of the whole grid (dimension 0 first).
  Tensors are ``torch`` tensors on ``device`` with shape
  ``(local_slices[, extent[1]], pitch)``; ``self.inputs`` / ``self.outputs`` hold
  the local arrays, ``self.own`` the local slice range this rank owns.
  """

  def __init__(self,
               program: launcher.CudaProgram,
               global_extent: Sequence[int],
               device: torch.device,
               rank: Optional[int] = None,
               world: Optional[int] = None,
               group=None,
               stream_handle: int = 0,
               exchange_every: Optional[int] = None):
    self.program = program
    self.group = group
    self.rank = dist.get_rank(group) if rank is None else rank
    self.world = dist.get_world_size(group) if world is None else world
    self.device = device
    self.global_extent = tuple(global_extent)
    self.dim = program.dim
    self.stream_handle = stream_handle
    s_dim = self.dim - 1
    total = self.global_extent[s_dim]
    self.ranges = split_slices(total, self.world)
    self.begin, self.end = self.ranges[self.rank]
    infos = [program.pass_info(i) for i in range(program.num_passes)]
    # clamped: a window strictly on one side of the stored cell reaches 0
    # slices the other way, never a negative number
    self.pass_reach = [(max(0, -info.reach_lo[s_dim]),
                        max(0, info.reach_hi[s_dim])) for info in infos]
    slab = min(end - begin for begin, end in self.ranges)
    self.groups = self._make_groups(exchange_every, slab)
    # ghost depth: what the deepest group needs before its first pass
    self.reach_lo = max(sum(self.pass_reach[i][0] for i in g)
                        for g in self.groups)
    self.reach_hi = max(sum(self.pass_reach[i][1] for i in g)
                        for g in self.groups)
    for begin, end in self.ranges:
      if end - begin < max(self.reach_lo, self.reach_hi):
        raise ValueError('slab thinner than the halo: use fewer ranks')
    # local array = owned slices + ghosts that exist globally
    self.local_begin = max(0, self.begin - self.reach_lo)
    self.local_end = min(total, self.end + self.reach_hi)
    self.local_extent = self.global_extent[:s_dim] + (self.local_end -
                                                     self.local_begin,)
    self.own = (self.begin - self.local_begin, self.end - self.local_begin)
    self.pitch = (self.global_extent[0] + 127) // 128 * 128
    self.inputs = [self._alloc(dt) for dt in program.input_dtypes]
    self.outputs = [self._alloc(dt) for dt in program.output_dtypes]
    n = len(program.output_dtypes)
    self.scratch = [[None] * n, [None] * n]
    self.launches = 0

  def _make_groups(self, exchange_every: Optional[int],
                   slab: int) -> List[List[int]]:
    """Consecutive passes that run between two halo exchanges."""
    n = self.program.num_passes
    if self.world == 1:
      return [list(range(n))]
    groups: List[List[int]] = []
    current: List[int] = []
    depth = 0
    budget = max(1, int(slab * 0.025))
    for index in range(n):
      reach = max(self.pass_reach[index])
      full = (len(current) >= exchange_every) if exchange_every else \
          (depth + reach > budget)
      if current and full:
        groups.append(current)
        current, depth = [], 0
      current.append(index)
      depth += reach
    groups.append(current)
    return groups

  # -- buffers -------------------------------------------------------------------
  def _torch_dtype(self, np_dtype):
    return getattr(torch, str(np_dtype))

  def _alloc(self, np_dtype) -> torch.Tensor:
    shape = tuple(self.local_extent[1:][::-1]) + (self.pitch,)
    return torch.zeros(shape, dtype=self._torch_dtype(np_dtype),
                       device=self.device)

  def _pitches(self):
    plane = self.pitch * self.local_extent[1] if self.dim == 3 else 0
    return (self.pitch, plane)

  def view(self, tensor: torch.Tensor) -> torch.Tensor:
    """The un-padded part of a local array."""
    return tensor[..., :self.global_extent[0]]

  # -- halo exchange ----------------------------------------------------------------
  def _comm_view(self, tensor: torch.Tensor) -> torch.Tensor:
    """NCCL (through torch) moves 8-bit, 32/64-bit signed integer and float
    types only - no int16, no unsigned 16/32/64 (found on B200 with the uint16
    blur: "Input tensor data type is not supported for NCCL process group:
    Short").  Halo slices of any other type travel as the same bytes."""
    native = (torch.int8, torch.uint8, torch.int32, torch.int64, torch.float16,
              torch.bfloat16, torch.float32, torch.float64)
    return tensor if tensor.dtype in native else tensor.view(torch.uint8)

  def start_exchange(self, tensors: Sequence[torch.Tensor], reach_lo: int,
                     reach_hi: int):
    """Starts filling the ghost slices of ``tensors`` from the neighbouring
    ranks and returns the pending work handles.

    A rank's lower ghost (``reach_lo`` slices) comes from the top of the rank
    below; its upper ghost (``reach_hi`` slices) from the bottom of the rank
    above.  The transfers are ordered after everything already queued on the
    current stream and run beside what is queued afterwards.
    """
    ops = []
    lo, hi = self.own
    for tensor in tensors:
      tensor = self._comm_view(tensor)
      if self.rank > 0:
        if reach_hi > 0:  # the lower neighbour's upper ghost is my bottom rows
          ops.append(dist.P2POp(dist.isend, tensor[lo:lo + reach_hi],
                                self._peer(self.rank - 1), self.group))
        if reach_lo > 0:
          ops.append(dist.P2POp(dist.irecv, tensor[lo - reach_lo:lo],
                                self._peer(self.rank - 1), self.group))
      if self.rank < self.world - 1:
        if reach_lo > 0:  # the upper neighbour's lower ghost is my top rows
          ops.append(dist.P2POp(dist.isend, tensor[hi - reach_lo:hi],
                                self._peer(self.rank + 1), self.group))
        if reach_hi > 0:
          ops.append(dist.P2POp(dist.irecv, tensor[hi:hi + reach_hi],
                                self._peer(self.rank + 1), self.group))
    return dist.batch_isend_irecv(ops) if ops else []

  def exchange(self, tensors: Sequence[torch.Tensor], reach_lo: int,
               reach_hi: int) -> None:
    for work in self.start_exchange(tensors, reach_lo, reach_hi):
      work.wait()

  def _peer(self, rank: int) -> int:
    if self.group is None:
      return rank
    return dist.get_global_rank(self.group, rank)

  # -- passes -------------------------------------------------------------------------
  def _boxes(self, last: bool):
    """Store boxes in local coordinates: this rank's own slices, clipped to the
    program's final valid box on the last pass."""
    prog = self.program
    s_dim = self.dim - 1
    lo_boxes, hi_boxes = [], []
    for o in range(len(prog.output_names)):
      lo = [0] * self.dim
      hi = list(self.local_extent)
      if last:
        final = prog.valid_box(o, self.global_extent)
        for d in range(s_dim):
          lo[d], hi[d] = final[d]
        g_lo = max(self.begin, final[s_dim][0])
        g_hi = min(self.end, final[s_dim][1])
      else:
        g_lo, g_hi = self.begin, self.end
      lo[s_dim] = g_lo - self.local_begin
      hi[s_dim] = max(lo[s_dim], g_hi - self.local_begin)
      lo_boxes.append(lo)
      hi_boxes.append(hi)
    return lo_boxes, hi_boxes

  def _launch(self, index, current, target, box_lo, box_hi, opts) -> None:
    pitches = self._pitches()
    self.program.run_pass(index, self.local_extent,
                          [t.data_ptr() for t in current],
                          [pitches] * len(current),
                          [t.data_ptr() for t in target],
                          [pitches] * len(target), box_lo, box_hi, opts)
    self.launches += 1

  def run(self, overlap: bool = True) -> None:
    """All ``iterate`` iterations: ``self.inputs`` -> ``self.outputs``.  The
    ghost slices of ``self.inputs`` are refreshed first, so callers only fill
    the slices they own.

    Passes run in exchange groups (module docstring).  Inside a group pass ``j``
    stores its own slices plus the ghost slices the later passes of the group
    still read; nobody is waited for.  With ``overlap`` the last pass of a group
    is issued as three launches: the slices next to each slab boundary (what
    the neighbours need for the next group) first, then the halo exchange of
    those slices is started, then the interior is computed while the exchange
    is in flight.
    """
    prog = self.program
    opts = launcher.make_opts(stream=self.stream_handle)
    s_dim = self.dim - 1
    total = self.global_extent[s_dim]
    own_lo, own_hi = self.own
    current = self.inputs

    def depth(group):
      return (sum(self.pass_reach[i][0] for i in group),
              sum(self.pass_reach[i][1] for i in group))

    self.exchange(current, *depth(self.groups[0]))
    for g, group in enumerate(self.groups):
      for k, index in enumerate(group):
        last = index == prog.num_passes - 1
        if last:
          target = self.outputs
        else:
          bank = self.scratch[index & 1]
          for o, dt in enumerate(prog.output_dtypes):
            if bank[o] is None:
              bank[o] = self._alloc(dt)
          target = bank
        box_lo, box_hi = self._boxes(last)
        # ghost slices the rest of the group still needs from this pass
        rest_lo = sum(self.pass_reach[i][0] for i in group[k + 1:])
        rest_hi = sum(self.pass_reach[i][1] for i in group[k + 1:])
        lo_slice = max(0, own_lo - rest_lo) if self.begin > 0 else own_lo
        hi_slice = min(self.local_extent[s_dim], own_hi + rest_hi) \
            if self.end < total else own_hi
        if not last:
          for o in range(len(box_lo)):
            box_lo[o][s_dim] = lo_slice
            box_hi[o][s_dim] = max(lo_slice, hi_slice)
        end_of_group = k == len(group) - 1
        if last or self.world == 1 or not end_of_group:
          self._launch(index, current, target, box_lo, box_hi, opts)
          current = target
          continue
        next_lo, next_hi = depth(self.groups[g + 1])
        # slices of this pass's output that a neighbour needs for the next group
        bottom = (own_lo, min(own_hi, own_lo + next_hi)) if self.rank > 0 \
            else (own_lo, own_lo)
        top = (max(bottom[1], own_hi - next_lo), own_hi) \
            if self.rank < self.world - 1 else (own_hi, own_hi)
        if not overlap:
          bottom, top = (own_lo, own_lo), (own_hi, own_hi)

        def restricted(lo_s, hi_s):
          lo = [list(b) for b in box_lo]
          hi = [list(b) for b in box_hi]
          for o in range(len(lo)):
            lo[o][s_dim] = max(lo[o][s_dim], lo_s)
            hi[o][s_dim] = max(lo[o][s_dim], min(hi[o][s_dim], hi_s))
          return lo, hi

        for lo_s, hi_s in (bottom, top):
          if hi_s > lo_s:
            self._launch(index, current, target, *restricted(lo_s, hi_s), opts)
        pending = self.start_exchange(target, next_lo, next_hi)
        if top[0] > bottom[1]:
          self._launch(index, current, target, *restricted(bottom[1], top[0]),
                       opts)
        for work in pending:
          work.wait()
        current = target
