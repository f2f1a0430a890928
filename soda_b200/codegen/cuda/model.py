"""Throughput model of a pass plan on B200: what bounds a (program, time
block) pair and which time block the planner should pick.

The reference sizes its accelerator the same way (reference:
src/soda/model/xilinx.py:131-144): performance = min(compute rate, DRAM
bandwidth / bytes per cell) x iterate.  Here, per pass variant:

  hbm    = HBM peak x eff / (bytes per cell per pass) x fused iterations
  fma    = fp32 lanes / (fp32 operations per update x halo redundancy)
  issue  = issue slots / (thread-instructions per update x halo redundancy)
  rate   = min(hbm, fma, issue)

* *halo redundancy*: cells a strip / tile computes per cell it stores; grows
  with the time block because the halo does (``strip / valid``), and jumps
  where the register windows force narrower lanes or fewer patch rows.
* *thread-instructions per update* are counted on the IR: arithmetic by
  operator (divisions, square roots and double precision weigh more), one
  shuffle per lane boundary a dimension-0 tap crosses, shared-memory traffic
  for 3-D dimension-1 taps outside the thread's patch, plus a per-step overhead
  (TMA ring, stores, window bookkeeping) that temporal blocking amortises.
  Packed fp32 / binary16 pairs halve the arithmetic issue slots, not the FMA
  pipe time (FADD2 holds the pipe two cycles).
* efficiencies are the fractions of each ceiling the shipped kernels reach on
  B200 (profiles/r01_ncu_full_*_summary.txt): 0.90 of the measured copy peak,
  0.68 of the FMA pipe, 0.70 of the issue slots.

The constants are validated against the measured sweep in
profiles/r02_time_block_sweep.jsonl (tests/test_model.py): the model only has
to rank time blocks, and its choice must be within a few per cent of the
measured best.
"""
from typing import Dict, List, Optional

from soda_b200 import ir, util
from soda_b200.codegen.cuda import plan as planner

HBM_GBS = 6535.1            # measured copy peak (MEASURED_PEAKS.json)
SM_COUNT = 148
LANES_PER_SM = 128
SM_GHZ = 1.965
LANE_RATE = SM_COUNT * LANES_PER_SM * SM_GHZ * 1e9  # thread-instr / s

HBM_EFF = 0.90
# fraction of the FMA pipe / issue slots the kernels reach.  Lanes of 8 cells
# carry twice the independent chains of 4-cell lanes at the same (low)
# occupancy: jacobi2d reaches 0.68 of the pipe at time block 6 (8 cells) and
# 0.60 at time block 8 (4 cells).  3-D CTAs meet at one barrier per plane; the
# more warps wait at it the fewer issue slots are used (jacobi3d: 0.61 with 4
# warps at time block 2, 0.49 with 12 at time block 3).
FMA_EFF = {8: 0.68, 4: 0.60}
ISSUE_EFF = 0.70
BARRIER_LOSS_PER_WARP = 0.025  # 3-D, per warp beyond 4
# thread-instructions per lane and step that do not depend on the program:
# input vector load from the TMA ring, store, pointer bumps, barrier / mbarrier
STEP_OVERHEAD_2D = 1.0   # per cell of the lane, per pass
STEP_OVERHEAD_3D = 1.5
LEVEL_OVERHEAD = 0.8     # per update: window rotation, predication, moves
MAX_REGISTERS = 255
REGISTER_SLACK = 52      # temporaries, addresses, store plans (tb6 kernel: 204
                         # registers for 152 window registers)


def _op_weights(node, counts: Dict[str, float]) -> None:
  """Accumulates weighted operation counts of an expression tree."""
  if isinstance(node, ir.Operand):
    _op_weights(node.inner, counts)
    return
  if isinstance(node, ir.BinaryOp):
    for operand in node.operand:
      _op_weights(operand, counts)
    if node.singleton:
      return
    t = ir.result_type(node)
    wide = 2.0 if t is not None and t.width_in_bits == 64 else 1.0
    for op in node.operator:
      if op == '/' or op == '%':
        if t is not None and t.is_float:
          counts['other'] += 9.0 * wide   # IEEE division: rcp + refinement
        else:
          counts['other'] += 4.0          # by a constant: IMAD.HI + fix-up
      elif t is not None and t.is_float and op in '+-*':
        counts['fp'] += wide
      else:
        counts['other'] += 1.0
    return
  if isinstance(node, ir.Unary):
    _op_weights(node.operand, counts)
    counts['other'] += 0.5 * len([op for op in node.operator if op != '+'])
    return
  if isinstance(node, ir.Cast):
    _op_weights(node.expr, counts)
    counts['other'] += 1.0
    return
  if isinstance(node, ir.Call):
    for arg in node.arg:
      _op_weights(arg, counts)
    if node.name == 'sqrt':
      counts['other'] += 12.0
    else:
      counts['other'] += max(1, len(node.arg) - 1)
    return


def stage_costs(stencil) -> List[Dict[str, float]]:
  """Per statement: weighted fp32-pipe and other operations per cell, and the
  distinct taps by kind."""
  costs = []
  for desc in planner.stage_descs(stencil):
    stmt = desc.stmt
    counts = {'fp': 0.0, 'other': 0.0}
    for let in stmt.let:
      _op_weights(let.expr, counts)
    _op_weights(stmt.expr, counts)
    taps = set()
    for ref in planner._stmt_loads(stmt, stencil.param_names):  # pylint: disable=protected-access
      taps.add((ref.name,) + tuple(a - b for a, b in zip(ref.idx,
                                                         stmt.ref.idx)))
    costs.append({'fp': counts['fp'], 'other': counts['other'], 'taps': taps})
  return costs


def estimate_pass(stencil, time_block: int,
                  options: Optional[Dict] = None,
                  extent: Optional[List[int]] = None) -> Optional[Dict]:
  """Model of one pass of ``time_block`` fused iterations with the planner's
  launch shape; ``None`` when the planner cannot build it."""
  try:
    pp = planner.make_tuned_pass_plan(stencil, time_block, options)
    if pp.dim == 3 and planner.smem_geometry_3d(pp)['bytes'] > \
        planner.SMEM_LIMIT_BYTES:
      return None
  except util.SemanticError:
    return None
  dim = pp.dim
  cells, cy = pp.cells, pp.cy
  window = sum(n.ring * max(1, n.haoda_type.width_in_bits // 32)
               for n in pp.nodes)
  registers = window * cells * cy + REGISTER_SLACK
  if registers > MAX_REGISTERS:
    return None

  # cells computed per cell stored
  if dim == 2:
    redundancy = pp.strip / pp.valid[0]
  else:
    redundancy = (pp.strip * pp.rows) / float(pp.valid[0] * pp.valid[1])
  if extent is not None:
    # whole strips / tiles: a 512-wide grid needs 5 strips of 120 valid cells
    tiles0 = -(-extent[0] // pp.valid[0])
    redundancy = tiles0 * pp.strip / float(extent[0])
    if dim == 3:
      tiles1 = -(-extent[1] // pp.valid[1])
      redundancy *= tiles1 * pp.rows / float(extent[1])

  fp = other = shuffles = smem = 0.0
  for cost in stage_costs(stencil):
    fp += cost['fp']
    other += cost['other']
    # one shuffle (packed: plus one move) per lane boundary a tap row crosses
    rows = {}
    for tap in cost['taps']:
      key = (tap[0],) + tap[2:]
      lo, hi = rows.get(key, (0, 0))
      rows[key] = (min(lo, tap[1]), max(hi, tap[1]))
    for key, (lo, hi) in rows.items():
      crossing = (-lo) + hi
      shuffles += crossing / float(cells) * (2.0 if pp.pack == 2 else 1.0)
      if dim == 3 and key[1] != 0:
        # dimension-1 neighbours outside the patch: one vector load per row
        # of the patch edge, and the producer's export of that row
        smem += min(1.0, abs(key[1]) / float(cy)) * 2.0 / 4.0
  arithmetic = (fp + other) / (2.0 if pp.pack == 2 else 1.0)
  overhead = (STEP_OVERHEAD_2D if dim == 2 else STEP_OVERHEAD_3D) / time_block
  instr = arithmetic + shuffles + smem + LEVEL_OVERHEAD + overhead

  bytes_per_cell = sum(t.width_in_bits // 8 for t in stencil.input_types +
                       stencil.output_types)
  hbm = HBM_GBS * 1e9 * HBM_EFF / bytes_per_cell * time_block
  issue_eff = ISSUE_EFF
  if dim == 3:
    warps = pp.rows // cy
    issue_eff *= max(0.4, 1.0 - BARRIER_LOSS_PER_WARP * max(0, warps - 4))
  fma_eff = FMA_EFF.get(cells, FMA_EFF[4] if cells < 8 else FMA_EFF[8])
  issue = LANE_RATE * issue_eff / (instr * redundancy)
  fma = LANE_RATE * fma_eff / (max(fp, 1e-9) * redundancy)
  rate = min(hbm, issue, fma)
  bound = 'hbm' if rate == hbm else ('issue' if rate == issue else 'fma')
  return {
      'time_block': time_block,
      'cells': cells,
      'cy': cy,
      'rows': pp.rows,
      'pack': pp.pack,
      'window_registers': window * cells * cy,
      'redundancy': redundancy,
      'instr_per_update': instr,
      'fp_per_update': fp,
      'hbm_ceiling_gcells': hbm / 1e9,
      'issue_ceiling_gcells': LANE_RATE / (instr * redundancy) / 1e9,
      'fma_ceiling_gcells': LANE_RATE / (max(fp, 1e-9) * redundancy) / 1e9,
      'gcells': rate / 1e9,
      'bound': bound,
  }


def estimate(stencil, time_block: int, options: Optional[Dict] = None,
             extent: Optional[List[int]] = None) -> Optional[Dict]:
  """Model of all ``iterate`` iterations run in passes of ``time_block`` (the
  remainder pass fuses fewer): cell updates per second over the schedule."""
  schedule = planner.pass_schedule(stencil.iterate, time_block)
  parts = {}
  seconds = 0.0
  for tb in set(schedule):
    part = estimate_pass(stencil, tb, options, extent)
    if part is None:
      return None
    parts[tb] = part
  for tb in schedule:
    seconds += tb / parts[tb]['gcells']
  main = dict(parts[schedule[0]])
  main['gcells'] = stencil.iterate / seconds
  main['passes'] = len(schedule)
  return main


# grid the launch shape is sized for when the caller gives no hint: the sizes
# of BASELINE.json's configs (whole strips / tiles matter: a 512-wide grid
# needs five 120-cell strips)
DEFAULT_EXTENT = {2: (16384, 16384), 3: (512, 512, 512)}


def choose_time_block(stencil, options: Optional[Dict] = None,
                      extent: Optional[List[int]] = None,
                      limit: int = 12) -> int:
  """The time block with the highest modelled throughput; ties and near-ties
  (within 2 %) go to the smaller one (shorter halos, shorter warm-up).
  ``extent`` (or ``options['extent_hint']``) is the grid the library will
  mostly run on; the program itself does not fix one."""
  extent = extent or (options or {}).get('extent_hint') or \
      DEFAULT_EXTENT.get(stencil.dim)
  best_tb, best = 1, None
  for tb in range(1, min(stencil.iterate, limit) + 1):
    est = estimate(stencil, tb, options, extent)
    if est is None:
      continue
    if best is None or est['gcells'] > best * 1.02:
      best_tb, best = tb, est['gcells']
  return best_tb
