"""Throughput model of a pass plan on B200: what bounds a (program, time
block) pair and which time block the planner should pick.

The reference sizes its accelerator the same way (reference:
src/soda/model/xilinx.py:131-144): performance = min(compute rate, DRAM
bandwidth / bytes per cell) x iterate.  Here, per pass variant and per cell:

  t_hbm   = bytes per cell per pass / HBM streaming rate
  t_fma   = fp32 operations per update x fused iterations x halo redundancy
            / (fp32 lanes x efficiency)
  t_issue = thread-instructions per update x fused iterations x halo
            redundancy / (issue slots x efficiency)
  t_pass  = (t_hbm^3 + t_fma^3 + t_issue^3)^(1/3)

The cubic norm instead of a plain maximum is what the measured sweep shows
(profiles/r02_time_block_sweep.jsonl): loads, arithmetic and stores of one
warp overlap only partly, so a pass slows down before it reaches either
ceiling (jacobi2d: 0.316 / 0.355 / 0.383 / 0.438 / 0.479 ms per pass at time
blocks 1 / 4 / 6 / 7 / 8; the norm gives 0.316 / 0.346 / 0.401 / 0.438 / 0.478).

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
  B200, fitted on the sweep: streaming stencils move 1.04 x the measured copy
  peak (less read/write turnaround than a copy); packed pairs keep the FMA
  pipe 0.72 busy with 8-cell lanes and 0.67 with 4-cell ones; scalar programs
  issue at 0.95 of the slots while at least three warps per scheduler are
  resident and at 0.67 below that (seidel2d from time block 5 on); 3-D CTAs
  issue at 0.80 with four-warp CTAs and 0.60 with larger ones (one CTA barrier
  per plane); every pass pays 8 us of launch.

The model only has to rank time blocks: tests/test_model.py checks that its
choice is within a few per cent of the measured best on every swept program.
"""
from typing import Dict, List, Optional

from soda_b200 import ir, util
from soda_b200.codegen.cuda import plan as planner

HBM_GBS = 6535.1            # measured copy peak (MEASURED_PEAKS.json)
SM_COUNT = 148
LANES_PER_SM = 128
SM_GHZ = 1.965
LANE_RATE = SM_COUNT * LANES_PER_SM * SM_GHZ * 1e9  # thread-instr / s

HBM_EFF = 1.04
FMA_EFF = {8: 0.72, 4: 0.67}
ISSUE_EFF = 0.95
ISSUE_EFF_LOW_OCCUPANCY = 0.67   # fewer than three warps per scheduler
# 3-D CTAs meet at one barrier per plane: four-warp CTAs, several per SM, hide
# it; the 12-16 warps of a time-block-3/4 tile wait for each other
ISSUE_EFF_3D = {True: 0.80, False: 0.60}   # keyed by (warps per CTA <= 4)
LAUNCH_SECONDS = 8e-6            # per pass: launch latency, ramp-up, tail
NORM = 3.0
NEAR_TIE = 0.12  # choose_time_block: how much throughput a smaller block may cost
REGISTER_FILE = 65536
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
  # resident warps per scheduler, from the register estimate
  threads = (tuning_warps(pp) or 4) * 32
  ctas = max(1, REGISTER_FILE // (min(registers, MAX_REGISTERS) * threads))
  warps_per_scheduler = ctas * threads / 32 / 4.0
  if dim == 3:
    issue_eff = ISSUE_EFF_3D[pp.rows // cy <= 4]
  elif pp.pack == 1 and warps_per_scheduler < 3.0:
    issue_eff = ISSUE_EFF_LOW_OCCUPANCY
  else:
    issue_eff = ISSUE_EFF
  fma_eff = FMA_EFF[8] if cells >= 8 else FMA_EFF[4]
  if pp.pack == 1:
    fma_eff = 1.0  # scalar FADD / FMUL: one pipe cycle each, the issue slots bind

  # seconds per grid cell and pass
  t_hbm = bytes_per_cell / (HBM_GBS * 1e9 * HBM_EFF)
  t_issue = instr * time_block * redundancy / (LANE_RATE * issue_eff)
  t_fma = fp * time_block * redundancy / (LANE_RATE * fma_eff)
  t_pass = (t_hbm ** NORM + t_issue ** NORM + t_fma ** NORM) ** (1.0 / NORM)
  if extent is not None:
    # small grids (C1: 2000 x 16384 cells take 30 us per pass) feel the launch
    grid_cells = 1
    for e in extent:
      grid_cells *= e
    t_pass += LAUNCH_SECONDS / grid_cells
  rate = time_block / t_pass
  bound = max((t_hbm, 'hbm'), (t_issue, 'issue'), (t_fma, 'fma'))[1]
  return {
      'time_block': time_block,
      'cells': cells,
      'cy': cy,
      'rows': pp.rows,
      'pack': pp.pack,
      'window_registers': window * cells * cy,
      'registers': registers,
      'warps_per_scheduler': warps_per_scheduler,
      'redundancy': redundancy,
      'instr_per_update': instr,
      'fp_per_update': fp,
      'hbm_ceiling_gcells': time_block / t_hbm / 1e9,
      'issue_ceiling_gcells': LANE_RATE / (instr * redundancy) / 1e9,
      'fma_ceiling_gcells': LANE_RATE / (max(fp, 1e-9) * redundancy) / 1e9,
      'ms_per_gcell_pass': t_pass * 1e12,
      'gcells': rate / 1e9,
      'bound': bound,
  }


def tuning_warps(pp) -> int:
  """Warps per CTA of a pass plan (the emitter's launch shape)."""
  from soda_b200.codegen.cuda import emit
  if pp.dim == 2:
    return emit.tuning_2d(pp, {})['kWarps']
  return pp.rows // pp.cy


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
  """The smallest time block whose modelled throughput is within 12 % of the
  best one.  Fusing more iterations than that buys little and costs halo
  (narrow grids, multi-GPU ghosts), warm-up slices per segment and registers
  (jacobi2d at time block 8 uses 252 of 255), and it moves the pass from the
  HBM side of the roofline to the FMA pipe, where a pass no longer streams at
  the memory system's speed: measured on B200, jacobi2d 16384^2 runs 4065
  Gcell-updates/s at 0.83 of the HBM roofline with time block 6 and 4442 at
  0.68 with time block 8 (profiles/r02_time_block_sweep.jsonl);
  ``--cuda-time-block 8`` asks for the latter.
  ``extent`` (or ``options['extent_hint']``) is the grid the library will
  mostly run on; the program itself does not fix one."""
  extent = extent or (options or {}).get('extent_hint') or \
      DEFAULT_EXTENT.get(stencil.dim)
  estimates = {}
  for tb in range(1, min(stencil.iterate, limit) + 1):
    est = estimate(stencil, tb, options, extent)
    if est is not None:
      estimates[tb] = est['gcells']
  if not estimates:
    return 1
  best = max(estimates.values())
  return min(tb for tb, gcells in estimates.items()
             if gcells >= (1.0 - NEAR_TIE) * best)
