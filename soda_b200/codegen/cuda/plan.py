"""Tiling / temporal-blocking planner over the stencil IR.

Turns one ``Stencil`` into a *pass plan*: the DAG of one HBM round trip, i.e.
``time_block`` consecutive iterations of the program's statements fused into
one kernel, together with everything the hand-written kernel templates need as
compile-time tables:

* ``lag``      - node ``n`` produces slice ``t - lag[n]`` of the streamed (last)
                 dimension at step ``t`` (the step at which input slice ``t``
                 arrives).  This is the closed form of the produce offsets that
                 the reference obtains from an ILP over stream offsets
                 (reference: src/soda/core.py:371-446).
* ``ring``     - how many slices of a node each thread keeps in its register
                 sliding window; the GPU counterpart of the reference's reuse
                 buffers, whose length is the reuse distance
                 (reference: src/soda/core.py:505-552,684-795).
* halos        - how far validity shrinks in the non-streamed dimensions after
                 the whole pass, which fixes the overlap between neighbouring
                 tiles (the reference's host tiling uses the stencil window the
                 same way, reference: src/soda/codegen/frt/host.py:124-131).

Streaming layout: threads of a warp own ``cells`` adjacent dimension-0 cells
each (so a warp covers ``32 * cells`` contiguous cells); the last dimension is
streamed slice by slice.  In 2-D every warp is an independent strip; in 3-D the
warps of a CTA are the rows (dimension 1) of a tile and exchange dimension-1
neighbours through shared-memory planes.
"""
import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

from soda_b200 import ir, util, visitor

WARP = 32


@dataclasses.dataclass
class Node:
  id: int
  kind: str  # 'input' | 'stage'
  src: int  # input index, or statement index within one iteration
  iteration: int
  name: str
  haoda_type: ir.Type
  prods: List[int] = dataclasses.field(default_factory=list)  # by slot K
  deltas: List[List[Tuple[int, ...]]] = dataclasses.field(
      default_factory=list)  # by slot K
  lag: int = 0
  ring: int = 1
  out: int = -1  # index of the output array this node is stored to, or -1
  # cells of the strip/tile that are invalid at the low / high end, per
  # non-streamed dimension
  halo_lo: Tuple[int, ...] = ()
  halo_hi: Tuple[int, ...] = ()
  # bounds of the dependency cone towards the pass inputs, per dimension
  win_lo: Tuple[int, ...] = ()
  win_hi: Tuple[int, ...] = ()
  # 3-D only: planes kept in shared memory for dimension-1 neighbours, and
  # the furthest dimension-1 offset at which any consumer reads them
  smem_depth: int = 0
  smem_reach: int = 0


@dataclasses.dataclass
class StageDesc:
  """One statement of the program (shared by all iterations of a pass)."""
  index: int
  stmt: object
  slots: List[str]  # distinct loaded tensor names, slot order


@dataclasses.dataclass
class PassPlan:
  dim: int
  time_block: int
  cells: int  # cells per lane (C)
  strip: int  # cells per warp in dimension 0 (32 * C)
  nodes: List[Node]
  stages: List[StageDesc]
  num_inputs: int
  num_outputs: int
  halo_lo: Tuple[int, ...]  # aligned halos of the pass outputs
  halo_hi: Tuple[int, ...]
  valid: Tuple[int, ...]  # valid cells per non-streamed dim per strip/tile
  lo_s: int  # lowest input slice (relative) an output slice depends on
  max_lag: int
  rows: int = 1  # 3-D: tile rows (= warps per CTA * rows per thread)
  cy: int = 1  # 3-D: rows of the patch one thread owns
  align0: int = 1  # strip origins are multiples of this many cells
  pack: int = 1  # cells evaluated per instruction (2: packed fp32 pairs)
  # 1: every node reads what its producers had produced *before* the current
  # step (software pipelining across the DAG), 0: producers run first
  skew: int = 0

  @property
  def output_nodes(self) -> List[Node]:
    return [n for n in self.nodes if n.out >= 0]


def stage_descs(stencil) -> List[StageDesc]:
  """Statements of one iteration in dependency order with their load slots."""
  stmts = list(stencil.local_stmts) + list(stencil.output_stmts)
  names = {s.name for s in stmts}
  deps = {}
  for s in stmts:
    loaded = set()
    for ref in _stmt_loads(s, stencil.param_names):
      if ref.name in names:
        loaded.add(ref.name)
    deps[s.name] = loaded
  order: List = []
  done = set()
  pending = list(stmts)
  while pending:
    progress = False
    for s in list(pending):
      if deps[s.name] <= done:
        order.append(s)
        done.add(s.name)
        pending.remove(s)
        progress = True
    if not progress:
      raise util.SemanticError('cyclic dependency among statements')
  result = []
  for i, s in enumerate(order):
    slots: List[str] = []
    for ref in _stmt_loads(s, stencil.param_names):
      if ref.name not in slots:
        slots.append(ref.name)
    result.append(StageDesc(index=i, stmt=s, slots=slots))
  return result


def _stmt_loads(stmt, params: Sequence[str] = ()) -> Tuple[ir.Ref, ...]:
  """Tensor loads of a statement; references to ``param`` arrays are constants,
  not loads (reference: src/soda/core.py:292-293)."""
  loads: Tuple[ir.Ref, ...] = ()
  for let in stmt.let:
    loads += visitor.get_load_tuple(let)
  loads += visitor.get_load_tuple(stmt.expr)
  return tuple(ref for ref in loads if ref.name not in params)


def _round_up(value: int, multiple: int) -> int:
  return (value + multiple - 1) // multiple * multiple


def default_cells(stencil) -> int:
  """Cells per lane: 16-byte vectors for the widest element type."""
  widest = max(t.width_in_bits for t in stencil.input_types +
               stencil.output_types + tuple(stencil.local_types))
  return max(2, 128 // max(widest, 16))


def packable(stencil) -> bool:
  """Whether the program can be evaluated two cells at a time with packed
  instructions (fp32: FADD2 / FFMA2, half: HADD2 / HMUL2): every tensor has the
  same type, ``float`` or ``half``, and every statement only adds, subtracts
  and multiplies loads, literals of that type (for half: a ``half(...)`` cast
  of a literal) and integer literals.  Anything else (division, calls,
  comparisons, wider literals - a float literal makes a half operation a float
  one -, integer tensors) keeps the scalar path.  A pair holds cells
  ``(u, u + C/2)`` of a lane, so a dimension-0 offset costs one shuffle and one
  move per lane boundary crossed, whatever its parity (soda_stream.cuh)."""
  types = stencil.input_types + stencil.output_types + tuple(
      stencil.local_types)
  elem_t = types[0]
  if elem_t not in (ir.FLOAT, ir.HALF) or any(t != elem_t for t in types):
    return False

  def literal(node) -> bool:
    while isinstance(node, ir.Operand) or (isinstance(node, ir.BinaryOp) and
                                           node.singleton):
      node = node.inner if isinstance(node, ir.Operand) else node.operand[0]
    return isinstance(node, ir.Num)

  def ok(node) -> bool:
    if isinstance(node, ir.Operand):
      return ok(node.inner)
    if isinstance(node, (ir.Ref, ir.Var)):
      return True
    if isinstance(node, ir.Num):
      return node.literal_type in (elem_t, ir.INT32)
    if isinstance(node, ir.Cast):
      # a cast of a literal is a scalar of the element type, broadcast
      return node.haoda_type == elem_t and (literal(node.expr) or
                                            ok(node.expr))
    if isinstance(node, ir.Unary):
      return all(op in '+-' for op in node.operator) and ok(node.operand)
    if isinstance(node, ir.AddSub):
      return all(ok(o) for o in node.operand)
    if isinstance(node, ir.MulDiv):
      return all(op == '*' for op in node.operator) and \
          all(ok(o) for o in node.operand)
    if isinstance(node, ir.BinaryOp) and node.singleton:
      return ok(node.operand[0])
    return False

  float_t = elem_t
  for stmt in list(stencil.local_stmts) + list(stencil.output_stmts):
    for let in stmt.let:
      if let.haoda_type not in (None, float_t) or not ok(let.expr):
        return False
    if not ok(stmt.expr):
      return False
    # every load at a dimension-0 offset costs a shuffle and a move per lane
    # to rotate a pair across the lane boundary; measured on B200, packing
    # pays while a statement has at most two of them (5/7-point stars: +8 % at
    # time block 6; the 9-point box of seidel2d with six: -5 %)
    shifted = {
        tuple(a - b for a, b in zip(ref.idx, stmt.ref.idx))
        for ref in _stmt_loads(stmt, stencil.param_names)
        if ref.idx[0] != stmt.ref.idx[0]
    }
    if len(shifted) > 2:
      return False
  return True


def make_pass_plan(stencil,
                   time_block: int = 1,
                   cells: Optional[int] = None,
                   rows: int = 8,
                   cy: int = 1,
                   pack: Optional[bool] = None,
                   pipelined: Optional[bool] = None) -> PassPlan:
  """Plans one pass of ``time_block`` fused iterations.

  ``rows`` and ``cy`` are only used by 3-D programs: a CTA covers ``rows``
  tile rows with ``rows / cy`` warps, every thread owning ``cy`` rows.

  ``pipelined`` (2-D, default on): node ``n`` only reads slices its producers
  finished in an earlier step, so the nodes of one step do not depend on each
  other and the ``time_block`` dependent chains of a step (5 additions each in
  jacobi2d) overlap instead of running back to back.  Costs one step of lag
  per DAG level and no registers: a window of depth ``d`` still holds the
  ``d`` slices a consumer reads, they are just one step older.
  """
  for stmt in stencil.param_stmts:
    if not stmt.haoda_type.is_executable or not stmt.size:
      raise util.SemanticError('param %s: unsupported type or shape' %
                               stmt.name)
  dim = stencil.dim
  if dim not in (2, 3):
    raise util.SemanticError(
        'the CUDA kernel templates take 2-D and 3-D programs, got %d-D '
        '(1-D programs are lifted to N x 1 by emit_program; 4-D ones are not '
        'supported)' % dim)
  if time_block < 1:
    raise util.SemanticError('time block must be positive')
  if time_block > 1 and len(stencil.input_stmts) != len(stencil.output_stmts):
    raise util.SemanticError('cannot fuse iterations of a program whose '
                             'inputs and outputs differ in number')
  for t in stencil.input_types + stencil.output_types + tuple(
      stencil.local_types):
    if not t.is_executable:
      raise util.SemanticError(
          'type %s is not supported by the CUDA backend%s' %
          (t, ' as such: lower it with soda_b200.optimization.widths.lower '
           '(emit_program does)' if t.is_lowerable else ''))
  if pipelined is None:
    pipelined = dim == 2
  if pipelined and dim != 2:
    raise util.SemanticError('the pipelined schedule is a 2-D feature')
  step_skew = 1 if pipelined else 0
  cells = cells or default_cells(stencil)
  strip = WARP * cells
  stages = stage_descs(stencil)
  num_inputs = len(stencil.input_stmts)
  num_outputs = len(stencil.output_stmts)
  input_names = stencil.input_names
  output_names = stencil.output_names
  s_dim = dim - 1

  nodes: List[Node] = []
  zero = (0,) * dim
  for i, stmt in enumerate(stencil.input_stmts):
    nodes.append(
        Node(id=i,
             kind='input',
             src=i,
             iteration=0,
             name=stmt.name,
             haoda_type=stmt.haoda_type,
             halo_lo=(0,) * (dim - 1),
             halo_hi=(0,) * (dim - 1),
             win_lo=zero,
             win_hi=zero))

  # name -> node id within the current iteration
  for iteration in range(time_block):
    table: Dict[str, int] = {}
    for i, name in enumerate(input_names):
      if iteration == 0:
        table[name] = i
      else:
        prev_out = output_names[i]
        table[name] = prev_iter_table[prev_out]  # noqa: F821
    for desc in stages:
      stmt = desc.stmt
      node = Node(id=len(nodes),
                  kind='stage',
                  src=desc.index,
                  iteration=iteration,
                  name=stencil.name_in_iter(stmt.name, iteration)
                  if stencil.iterate > iteration else stmt.name,
                  haoda_type=stmt.haoda_type)
      by_slot: Dict[str, List[Tuple[int, ...]]] = {n: [] for n in desc.slots}
      for ref in _stmt_loads(stmt, stencil.param_names):
        delta = tuple(a - b for a, b in zip(ref.idx, stmt.ref.idx))
        if delta not in by_slot[ref.name]:
          by_slot[ref.name].append(delta)
      for name in desc.slots:
        node.prods.append(table[name])
        node.deltas.append(by_slot[name])
      if iteration == time_block - 1 and stmt.name in output_names:
        node.out = output_names.index(stmt.name)
      nodes.append(node)
      table[stmt.name] = node.id
    prev_iter_table = table  # noqa: F841

  # lags, window bounds and halos along the DAG (nodes are in dependency order)
  for node in nodes:
    if node.kind == 'input':
      continue
    lag = 0
    win_lo: List[Optional[int]] = [None] * dim
    win_hi: List[Optional[int]] = [None] * dim
    halo_lo = [0] * (dim - 1)
    halo_hi = [0] * (dim - 1)
    for prod_id, deltas in zip(node.prods, node.deltas):
      prod = nodes[prod_id]
      for delta in deltas:
        skew = step_skew
        if dim == 3 and delta[1] != 0 and prod.kind != 'input':
          # dimension-1 neighbours come from shared memory written by other
          # warps in an earlier step: read one step late, one barrier per step
          skew = 1
        lag = max(lag, prod.lag + max(delta[s_dim], 0) + skew,
                  prod.lag + delta[s_dim] + skew)
        for d in range(dim):
          lo, hi = prod.win_lo[d] + delta[d], prod.win_hi[d] + delta[d]
          win_lo[d] = lo if win_lo[d] is None else min(win_lo[d], lo)
          win_hi[d] = hi if win_hi[d] is None else max(win_hi[d], hi)
        for d in range(dim - 1):
          halo_lo[d] = max(halo_lo[d], prod.halo_lo[d] - delta[d])
          halo_hi[d] = max(halo_hi[d], prod.halo_hi[d] + delta[d])
    node.lag = lag
    node.win_lo = tuple(0 if v is None else v for v in win_lo)
    node.win_hi = tuple(0 if v is None else v for v in win_hi)
    node.halo_lo = tuple(halo_lo)
    node.halo_hi = tuple(halo_hi)

  # The lags above are as-soon-as-possible.  A stage whose consumers all run
  # later would only sit in a register window until then, so every stage that
  # is not stored moves as late as its consumers allow (windows get shorter,
  # fewer registers, fewer moves per step).  Stored nodes keep their lag: the
  # pass latency does not grow.
  def edge_skew(consumer: Node, prod: Node, delta) -> int:
    if dim == 3 and delta[1] != 0 and prod.kind != 'input':
      return 1
    return step_skew

  for node in reversed(nodes):
    if node.kind == 'input' or node.out >= 0:
      continue
    latest = None
    for consumer in nodes:
      for prod_id, deltas in zip(consumer.prods, consumer.deltas):
        if prod_id != node.id:
          continue
        for delta in deltas:
          bound = consumer.lag - max(delta[s_dim], 0) - edge_skew(
              consumer, node, delta)
          latest = bound if latest is None else min(latest, bound)
    if latest is not None and latest > node.lag:
      node.lag = latest

  # register rings and shared-memory plane depths
  for node in nodes:
    ring = 1
    smem_depth = 0
    smem_reach = 0
    for consumer in nodes:
      for prod_id, deltas in zip(consumer.prods, consumer.deltas):
        if prod_id != node.id:
          continue
        for delta in deltas:
          distance = consumer.lag - node.lag - delta[s_dim]
          if distance < 0:
            raise util.InternalError('negative reuse distance')
          if distance < step_skew:
            raise util.InternalError('consumer not behind its producer')
          if dim == 3 and delta[1] != 0:
            # rows outside the consumer's patch come from shared memory ...
            smem_depth = max(smem_depth, distance + 1)
            smem_reach = max(smem_reach, abs(delta[1]))
          if dim != 3 or delta[1] == 0 or cy > 1:
            # ... all others from the register window
            ring = max(ring, distance + 1 - step_skew)
    node.ring = ring
    node.smem_depth = smem_depth
    node.smem_reach = smem_reach

  outs = [n for n in nodes if n.out >= 0]
  halo_lo = [max(n.halo_lo[d] for n in outs) for d in range(dim - 1)]
  halo_hi = [max(n.halo_hi[d] for n in outs) for d in range(dim - 1)]
  # dimension 0: strip origins must be multiples of the lane vector (aligned
  # vector stores) and of 16 bytes of every input (TMA faults on a box whose
  # first element is not 16-byte aligned in global memory)
  # (whole lanes are stored or not: halos and valid widths are multiples of
  # the lane width; 6-cell lanes, which would cover a 512-wide grid with three
  # strips instead of five, fail on exactly this: lcm(6 cells, 16 bytes) = 12
  # cells of low halo eat the gain)
  align0 = max([cells] + [
      128 // t.width_in_bits for t in stencil.input_types
  ])
  halo_lo[0] = _round_up(halo_lo[0], align0)
  valid0 = (strip - halo_lo[0] - halo_hi[0]) // align0 * align0
  if valid0 <= 0:
    raise util.SemanticError(
        'stencil window (%d cells in dimension 0 after %d fused iterations) '
        'does not fit a %d-cell strip; lower the time block' %
        (halo_lo[0] + halo_hi[0], time_block, strip))
  valid = [valid0]
  if dim == 3:
    if cy < 1 or rows % cy:
      raise util.SemanticError(
          'tile rows (%d) must be a multiple of the rows per thread (%d)' %
          (rows, cy))
    valid1 = rows - halo_lo[1] - halo_hi[1]
    if valid1 <= 0:
      raise util.SemanticError(
          'stencil window does not fit a %d-row tile; lower the time block or '
          'raise the tile rows' % rows)
    valid.append(valid1)

  return PassPlan(dim=dim,
                  time_block=time_block,
                  cells=cells,
                  strip=strip,
                  nodes=nodes,
                  stages=stages,
                  num_inputs=num_inputs,
                  num_outputs=num_outputs,
                  halo_lo=tuple(halo_lo),
                  halo_hi=tuple(halo_hi),
                  valid=tuple(valid),
                  lo_s=min(n.win_lo[s_dim] for n in outs),
                  max_lag=max(n.lag for n in outs),
                  rows=rows if dim == 3 else 1,
                  cy=cy if dim == 3 else 1,
                  align0=align0,
                  pack=2 if (pack is not False and cells % 2 == 0 and
                             packable(stencil)) else 1,
                  skew=step_skew)


# dynamic shared memory one CTA may ask for on sm_100a (227 KB)
SMEM_LIMIT_BYTES = 227 * 1024
MAX_CTA_THREADS = 1024


def smem_geometry_3d(pass_plan: PassPlan, lookahead: int = 2) -> Dict[str, int]:
  """Shared-memory footprint of a 3-D pass, the arithmetic of Smem3D in
  soda_stream.cuh: a TMA ring of ``stages`` slots (one plane per input each),
  ``smem_depth`` exported planes per fused stage that is read across tile rows,
  guard bytes on both sides, the barriers."""
  in_depth = max(
      [1] + [n.smem_depth for n in pass_plan.nodes if n.kind == 'input'])
  reach = 0
  elem = 4
  for node in pass_plan.nodes:
    elem = max(elem, node.haoda_type.width_in_bits // 8)
    for deltas in node.deltas:
      for delta in deltas:
        if delta[1] != 0:
          reach = max(reach,
                      abs(delta[1]) * pass_plan.strip + abs(delta[0]) +
                      pass_plan.cells)
  guard = (reach * elem + 127) // 128 * 128
  stages = in_depth + lookahead
  plane_cells = pass_plan.strip * pass_plan.rows
  slot = sum(n.haoda_type.width_in_bits // 8 * plane_cells
             for n in pass_plan.nodes if n.kind == 'input')
  exports = sum(n.smem_depth * (n.haoda_type.width_in_bits // 8) * plane_cells
                for n in pass_plan.nodes if n.kind == 'stage')
  barriers = (guard + slot * stages + exports + guard + 127) // 128 * 128
  return {'in_depth': in_depth, 'guard': guard, 'stages': stages,
          'barrier_offset': barriers, 'bytes': barriers + 8 * stages}


def make_tuned_pass_plan(stencil, time_block: int,
                         options: Optional[Dict] = None) -> PassPlan:
  """``make_pass_plan`` with the launch shape chosen from the DAG.

  Explicit ``options`` (cells, rows, cy, no_pack, no_pipeline) win.  The
  defaults follow what was measured on B200:

  * 2-D: 8 fp32 cells per lane instead of 4 (half the strip overlap and half
    the shuffles per cell) while the register windows stay below ~160
    registers; pipelined schedule.
  * 3-D: unpacked arithmetic; as many patch rows per thread (4, 2, 1) as keep
    the windows below ~120 registers; tile rows = 4 x the dimension-1 halo,
    at least 8 - small CTAs, several per SM, hide the per-step barrier better
    than one large CTA; 6 x the halo for DAGs too large for patches.
  """
  options = dict(options or {})
  dim = stencil.dim
  pack = False if options.get('no_pack') else None
  pipelined = False if options.get('no_pipeline') else None
  cells = options.get('cells')
  probe = make_pass_plan(stencil, time_block=time_block, cells=cells,
                         rows=256, cy=1, pack=pack, pipelined=pipelined)
  window = sum(n.ring * max(1, n.haoda_type.width_in_bits // 32)
               for n in probe.nodes)
  if dim == 2:
    # 8- and 16-bit cells stay at 8 per lane.  16-cell lanes work (`cells`
    # option; the 512-cell strip then arrives as two TMA boxes) but measured
    # slower on B200 for every program tried: blur time block 2 1577 vs 1791
    # Gcell-updates/s, sobel2d 1382 vs 1451, half jacobi2d time block 6 4368
    # vs 4826 (profiles/r01_lanes16.jsonl)
    # register budget of 8-cell lanes: packed pair arithmetic needs few
    # temporaries (jacobi2d time block 8: 200 window registers, 252 in all, no
    # spill, 4483 against 3941 Gcell-updates/s with 4-cell lanes); scalar
    # arithmetic slows down well before it spills (seidel2d time block 5: 128
    # window registers, 199 in all, 0.64 ms per pass against 0.41 at time
    # block 4 - profiles/r02_time_block_sweep.jsonl)
    budget = 200 if probe.pack == 2 else 160
    if cells is None and probe.cells * 2 * window <= budget and all(
        t.width_in_bits == 32 for t in stencil.input_types +
        stencil.output_types + tuple(stencil.local_types)):
      cells = probe.cells * 2
      try:
        return make_pass_plan(stencil, time_block=time_block, cells=cells,
                              pack=pack, pipelined=pipelined)
      except util.SemanticError:
        cells = None
    return make_pass_plan(stencil, time_block=time_block, cells=cells,
                          pack=pack, pipelined=pipelined)
  if pack is None and not options.get('pack'):
    pack = False
  cy = options.get('cy')
  if not cy:
    cy = 1
    for candidate in (4, 2):
      if window * probe.cells * candidate <= 120:
        cy = candidate
        break
  rows = options.get('rows')
  if not rows:
    halo = probe.halo_lo[1] + probe.halo_hi[1]
    # large DAGs (one row per thread) are issue-bound: a taller tile wastes
    # less work on the dimension-1 halo (denoise3d: 24 rows 101, 16 rows 80
    # Gcell-updates/s)
    rows = max(8, (6 if cy == 1 else 4) * halo)
    rows = _round_up(rows, cy)
    # a CTA has rows / cy warps and its planes must fit shared memory: deep
    # multi-input DAGs (several exported stages, wide dimension-1 halos) get
    # a lower tile, down to one patch row more than the halo
    rows = min(rows, MAX_CTA_THREADS // 32 * cy)
    while rows - cy > halo:
      trial = make_pass_plan(stencil, time_block=time_block, cells=cells,
                             rows=rows, cy=cy, pack=pack, pipelined=pipelined)
      if smem_geometry_3d(trial)['bytes'] <= SMEM_LIMIT_BYTES:
        return trial
      rows -= cy
  return make_pass_plan(stencil, time_block=time_block, cells=cells, rows=rows,
                        cy=cy, pack=pack, pipelined=pipelined)


def choose_time_block(stencil, requested: Optional[int] = None,
                      options: Optional[Dict] = None) -> int:
  """Iterations fused per HBM round trip.  An explicit request wins; otherwise
  the throughput model (model.py: HBM, FMA-pipe and issue ceilings against the
  halo growth and the register windows of every candidate) picks it - the
  counterpart of the reference's performance model, which sizes ``iterate``
  the same way (reference: src/soda/model/xilinx.py:131-144)."""
  if stencil.iterate == 1 or len(stencil.input_stmts) != len(
      stencil.output_stmts):
    return 1
  if requested:
    return max(1, min(requested, stencil.iterate))
  from soda_b200.codegen.cuda import model  # model imports this module
  return model.choose_time_block(stencil, options)


def pass_schedule(iterate: int, time_block: int) -> List[int]:
  """Iterations fused by each successive pass, e.g. 10 by 4 -> [4, 4, 2]."""
  full, rest = divmod(iterate, time_block)
  return [time_block] * full + ([rest] if rest else [])
