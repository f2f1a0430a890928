#!/usr/bin/env python3
"""Benchmark harness: Gcell-updates/s of the SODA CUDA backend on B200.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Headline workload (BASELINE.json configs[1]): jacobi2d, fp32 5-point, 16384 x
16384, iterate 64, synthetic U[0,1) input; the planner picks the time block.
One "step" = all 64 iterations over the grid.  With N > 1 (launched by
torchrun, one rank per GPU) every rank owns a 16384 x 16384 slab of a 16384 x
(16384 N) grid (weak scaling) and swaps halo rows with its neighbours.

Prints ONE JSON line (see the driver contract) carrying, besides the metric:
  roofline      - per-pass HBM roofline of the dominant kernel, from CUDA events
  parity        - windows of the measured output, bit for bit against the g++
                  oracle run on their dependency cones (oracle/cone.py); at
                  N > 1 every rank checks windows next to its slab boundaries
  other_configs - BASELINE configs C1, C3 (heat3d, jacobi3d), C4 (denoise3d with
                  and without --computation-reuse) measured in the same process:
                  Gcell-updates/s, time block, per-pass roofline fraction, the
                  min(HBM, ALU) ceiling, parity windows
  c5_strong     - BASELINE config C5: jacobi2d 65536 x 65536, iterate 256, split
                  over the N GPUs of this run (strong scaling)
  cpu_baseline  - the g++-compiled restatement of the reference's loops (the
                  oracle, kind "port") timed on this box's host cores
  e2e           - the same metric through the public host-array API, with the
                  host->device and device->host copies inside the timed region,
                  next to the box's measured host<->device copy peak
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md
SM_COUNT = 148
LANE_INSTR_PER_S = SM_COUNT * 128 * 1.965e9  # thread-instructions / s at boost

# BASELINE.json configs: key -> program, Stencil overrides, grid.  Time block
# and launch shape are the planner's (soda_b200/codegen/cuda/plan.py).
CONFIGS = {
    'C1_blur': dict(program='blur', overrides={'iterate': 2},
                    extent=(2000, 16384)),
    'C2_jacobi2d': dict(program='jacobi2d', overrides={'iterate': 64},
                        extent=(16384, 16384)),
    # the headline program with 8 fused iterations (--cuda-time-block 8): 9 %
    # faster than the planner's 6, FMA-pipe-bound instead of HBM-bound
    'C2_jacobi2d_tb8': dict(program='jacobi2d', overrides={'iterate': 64},
                            extent=(16384, 16384), time_block=8),
    'C3_heat3d': dict(program='heat3d', overrides={'iterate': 32},
                      extent=(512, 512, 512)),
    # heat3d's coefficients are powers of two: with --cuda-pow2-fma every
    # `c * x + acc` is one FFMA (1 FMUL + 6 FFMA instead of 7 FMUL + 6 FADD per
    # update), exact unless c * x is subnormal (soda::fma_pow2)
    'C3_heat3d_pow2_fma': dict(program='heat3d', overrides={'iterate': 32},
                               extent=(512, 512, 512),
                               options={'pow2_fma': True}),
    'C3_jacobi3d': dict(program='jacobi3d', overrides={'iterate': 32},
                        extent=(512, 512, 512)),
    'C4_denoise3d': dict(program='denoise3d', overrides={},
                         extent=(512, 512, 512)),
    'C4_denoise3d_cr': dict(program='denoise3d',
                            overrides={'computation_reuse': 'yes'},
                            extent=(512, 512, 512)),
    # the same program with sqrt(float) kept in float (--math-precision float)
    'C4_denoise3d_float_math': dict(program='denoise3d',
                                    overrides={'math_precision': 'float'},
                                    extent=(512, 512, 512)),
    'C5_jacobi2d': dict(program='jacobi2d', overrides={'iterate': 256},
                        extent=(65536, 65536)),
}
# halo transport of the slab runtime at N > 1: 'nccl' = the library's own NCCL
# communicator (default), 'torch' = torch.distributed point-to-point callbacks
TRANSPORT = os.environ.get('SODA_BENCH_TRANSPORT', 'nccl')
HEADLINE = 'C2_jacobi2d'
OTHER = ('C1_blur', 'C2_jacobi2d_tb8', 'C3_heat3d', 'C3_heat3d_pow2_fma',
         'C3_jacobi3d', 'C4_denoise3d',
         'C4_denoise3d_cr', 'C4_denoise3d_float_math')
WIDTH, HEIGHT = CONFIGS[HEADLINE]['extent']
ITERATE = CONFIGS[HEADLINE]['overrides']['iterate']
PROGRAM = CONFIGS[HEADLINE]['program']


def config_stencil(key, **extra):
  from soda_b200 import sodac
  cfg = CONFIGS[key]
  with open(os.path.join(ROOT, 'tests', 'src', cfg['program'] + '.soda')) as fp:
    return sodac.compile_source(fp.read(), **dict(cfg['overrides'], **extra))


def config_program(key):
  """(Stencil, loaded CudaProgram) of a config, planner defaults."""
  from soda_b200.codegen import cuda as cuda_backend
  st = config_stencil(key)
  return st, cuda_backend.compile_stencil(
      st, time_block=CONFIGS[key].get('time_block'),
      options=CONFIGS[key].get('options'))


def measured_hbm_peak():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as fp:
      return float(json.load(fp)['hbm_gbs']), 'measured'
  except Exception:  # pylint: disable=broad-except
    return FALLBACK_HBM_GBS, 'fallback'


# ---------------------------------------------------------------------------
# synthetic data: a pure function of the global cell coordinates
# ---------------------------------------------------------------------------
def synthetic(box, seed, dtype, device):
  """Cells ``box`` (``(lo, hi)`` per dimension, dimension 0 first) of the
  synthetic input number ``seed`` as a torch tensor on ``device``.

  An integer hash of the global coordinates, scaled to U[0,1) for float
  tensors and masked for integer ones - integer operations and one exact
  conversion only, so the CPU (oracle cones, any rank) and the GPU (the
  measured arrays) produce the same bits for the same cell.
  """
  import torch
  dim = len(box)
  primes = (73856093, 19349663, 83492791)
  h = None
  for d in range(dim):
    lo, hi = box[d]
    coord = torch.arange(lo, hi, dtype=torch.int64, device=device)
    shape = [1] * dim
    shape[dim - 1 - d] = hi - lo
    term = ((coord * primes[d]) & 0x7FFFFFFF).reshape(shape)
    h = term if h is None else h ^ term
  h = h ^ ((seed + 1) * 40503)
  h = ((h ^ (h >> 13)) * 1274126177) & 0x7FFFFFFF
  h = h ^ (h >> 16)
  name = str(dtype).replace('torch.', '')
  if name.startswith('float'):
    return ((h & 0xFFFFFF).to(torch.float32) *
            (1.0 / 16777216.0)).to(getattr(torch, name))
  if name in ('uint16', 'uint32', 'uint64'):
    return (h & 0xFFFF).to(torch.int32).to(getattr(torch, name))
  if name in ('uint8',):
    return (h & 0xFF).to(torch.uint8)
  return ((h & 0x7FF) - 1024).to(getattr(torch, name))


def fill_synthetic(tensor, origin, seed, extent0):
  """Fills ``tensor`` (shape (..., pitch)) with the synthetic input whose cell
  (0, .., 0) of the last (streamed) axis sits at global slice ``origin``; in
  blocks, so that the int64 temporaries stay small."""
  dim = tensor.dim()
  slices = tensor.shape[0]
  rows = tensor.shape[1] if dim == 3 else 1
  block = max(1, (1 << 24) // (extent0 * rows))
  for s0 in range(0, slices, block):
    s1 = min(slices, s0 + block)
    box = [(0, extent0)]
    if dim == 3:
      box.append((0, rows))
    box.append((origin + s0, origin + s1))
    tensor[s0:s1, ..., :extent0].copy_(
        synthetic(tuple(box), seed, tensor.dtype, tensor.device))


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
  """Samples SM clocks and throttle reasons with nvidia-smi while running."""
  QUERY = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
           'clocks_event_reasons.hw_thermal_slowdown,'
           'clocks_event_reasons.sw_thermal_slowdown,'
           'clocks_event_reasons.sw_power_cap')

  def __init__(self, index=0):
    self.index = index
    self.samples = []
    self.proc = None
    self.thread = None

  def start(self):
    try:
      self.proc = subprocess.Popen([
          'nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
          '--format=csv,noheader,nounits', '-lms', '20'
      ], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
      return
    self.thread = threading.Thread(target=self._read, daemon=True)
    self.thread.start()

  def _read(self):
    for line in self.proc.stdout:
      self.samples.append((time.perf_counter(), line.strip()))

  def stop(self):
    if self.proc is not None:
      self.proc.terminate()
      try:
        self.proc.wait(timeout=5)
      except subprocess.TimeoutExpired:
        self.proc.kill()

  def summary(self, t0, t1):
    sm, sm_max, reasons = [], 0, set()
    names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
             'sw_power_cap']
    rows = [s for t, s in self.samples if t0 <= t <= t1] or \
        [s for _, s in self.samples[-3:]]
    for row in rows:
      parts = [p.strip() for p in row.split(',')]
      try:
        sm.append(float(parts[0]))
        sm_max = max(sm_max, float(parts[1]))
      except (ValueError, IndexError):
        continue
      for name, flag in zip(names, parts[2:]):
        if flag.lower().startswith('active'):
          reasons.add(name)
    sm.sort()
    return {
        'sm_mhz': sm[len(sm) // 2] if sm else None,
        'sm_max_mhz': sm_max or None,
        'reasons': sorted(reasons),
        'samples': len(sm),
    }


# ---------------------------------------------------------------------------
# CPU baseline (the oracle's g++ build)
# ---------------------------------------------------------------------------
_CPU_STATE = {}


def cpu_baseline(sample_iterate=4, repeats=2):
  import numpy as np
  from oracle import emit_cpp
  if sample_iterate not in _CPU_STATE:
    st = config_stencil(HEADLINE, iterate=sample_iterate)
    rng = np.random.default_rng(1)
    grid = rng.random((HEIGHT, WIDTH), dtype=np.float32)
    _CPU_STATE[sample_iterate] = (emit_cpp.Oracle(st, timed=True), grid,
                                  {'t0': np.zeros_like(grid)})
  oracle, grid, out = _CPU_STATE[sample_iterate]
  best = None
  for _ in range(repeats):
    t0 = time.perf_counter()
    oracle.run({'t1': grid}, out)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
  cores = os.cpu_count() or 1
  threads = int(os.environ.get('OMP_NUM_THREADS', cores))
  return {
      'value': WIDTH * HEIGHT * sample_iterate / best / 1e9,
      'unit': 'Gcell-updates/s',
      'cores': threads,
      'kind': 'port',
      'sample': '%dx%d grid, %d of %d iterations, best of %d, g++ -O3 '
                '-march=native -fopenmp -ffp-contract=off' %
                (WIDTH, HEIGHT, sample_iterate, ITERATE, repeats),
      'seconds': best,
  }


DATAFLOW_SAMPLE = dict(tile=2048, rows=1024, iterate=8)


def dataflow_stencil():
  """The headline program as the reference's FPGA flow would see a bounded
  sample of it: one 2048-wide tile, unroll factor 2, 8 iterations."""
  return config_stencil(HEADLINE, iterate=DATAFLOW_SAMPLE['iterate'],
                        tile_size=[DATAFLOW_SAMPLE['tile']], unroll_factor=2)


def cpu_baseline_dataflow():
  """Second CPU baseline: the reference-style
  dataflow kernel - modules, FIFOs, delay lines, burst words, printed by
  oracle/dataflow_kernel.py - under C-simulation semantics (modules run one
  after the other over unbounded streams, single-threaded by construction)."""
  import numpy as np
  from oracle import dataflow_kernel
  from soda_b200.codegen.cuda import stream_layout
  st = dataflow_stencil()
  extent = (DATAFLOW_SAMPLE['tile'], DATAFLOW_SAMPLE['rows'])
  kernel = dataflow_kernel.DataflowKernel(st, timed=True)
  layout = stream_layout.TensorLayout(st, st.input_names[0], extent)
  cycles = dataflow_kernel.cycle_count(st, extent)
  rng = np.random.default_rng(2)
  in_banks = {st.input_names[0]: [rng.random(layout.elems_per_bank,
                                             dtype=np.float32)]}
  out_banks = {st.output_names[0]: [np.zeros(layout.elems_per_bank,
                                             dtype=np.float32)]}
  t0 = time.perf_counter()
  kernel.run(in_banks, out_banks, cycles)
  seconds = time.perf_counter() - t0
  counts = dataflow_kernel.summary(st)
  return {
      'value': extent[0] * extent[1] * st.iterate / seconds / 1e9,
      'unit': 'Gcell-updates/s',
      'cores': 1,
      'kind': 'port (reference-style dataflow kernel, C simulation)',
      'sample': '%dx%d grid as one tile, %d of %d iterations, unroll factor 2, '
                'kernel only (streams already tiled), g++ -O3' %
                (extent[0], extent[1], st.iterate, ITERATE),
      'modules': {k: counts[k] for k in ('load', 'forward', 'compute', 'store')},
      'fifos': counts['fifos'],
      'seconds': seconds,
  }


def workload_config(n_gpus, time_block=None, passes=None):
  return {
      'workload': 'jacobi2d fp32 5-point %dx%d iterate %d' %
                  (WIDTH, HEIGHT * n_gpus, ITERATE),
      'program': 'tests/src/jacobi2d.soda --iterate %d' % ITERATE,
      'grid_per_gpu': [WIDTH, HEIGHT],
      'time_block': time_block,
      'time_block_chosen_by': 'planner (soda_b200/codegen/cuda/model.py)',
      'passes': passes,
      'parallelism': 'slab%d' % n_gpus if n_gpus > 1 else 'single',
      'l2': 'inputs (1 GiB per array per GPU) are larger than the 126 MB L2',
  }


def run_reference(args, out):
  """--impl reference: the reference's CPU path (its golden loops, restated and
  compiled with g++ because the reference generator cannot run here) on all
  host threads.  Rank 0 only."""
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  # torchrun pins OMP_NUM_THREADS=1 per rank; the baseline uses every core
  os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)
  sample_iterate = 4
  steps = max(1, args.steps)
  baseline = None
  t_total = 0.0
  for _ in range(max(0, min(args.warmup, 1))):
    cpu_baseline(sample_iterate, repeats=1)
  for _ in range(steps):
    baseline = cpu_baseline(sample_iterate, repeats=1)
    t_total += baseline['seconds']
  value = WIDTH * HEIGHT * sample_iterate * steps / t_total / 1e9
  baseline['value'] = value
  line = {
      'impl': 'reference',
      'metric': 'Gcell-updates/s',
      'value': value,
      'unit': 'Gcell-updates/s',
      'n_gpus': args.gpus,
      'steps': steps,
      'warmup': min(args.warmup, 1),
      'ms_per_step': t_total / steps * 1e3,
      'higher_is_better': True,
      'scaling': 'weak',
      'vs_baseline': None,
      'dtype': 'f32',
      'data': 'synthetic',
      'config': workload_config(1),
      'cpu_baseline': baseline,
      'e2e': {'value': value, 'unit': 'Gcell-updates/s',
              'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }
  out.emit(json.dumps(line))


# ---------------------------------------------------------------------------
# parity against the oracle, by dependency cones
# ---------------------------------------------------------------------------
def cone_parity(st, extent, read_output, count, seed=0, required=(),
                rows=None):
  """Windows of the measured output against the g++ oracle on their cones.
  Inputs are regenerated on the CPU from the synthetic hash.  ``rows``: the
  streamed-dimension range (global) this caller can read outputs of."""
  import numpy as np
  import torch
  from oracle import cone
  dim = st.dim
  size = (64, 64) if dim == 2 else (24, 16, 16)
  valid = [list(b) for b in st.valid_box(st.output_names[0], extent)]
  if rows is not None:
    valid[dim - 1] = [max(valid[dim - 1][0], rows[0]),
                      min(valid[dim - 1][1], rows[1])]
  valid = tuple(tuple(b) for b in valid)
  windows = cone.draw_windows(valid, size, count, seed, required)
  cpu = torch.device('cpu')
  dtypes = {}
  for index, stmt in enumerate(st.input_stmts):
    from oracle import golden
    dtypes[stmt.name] = (index, getattr(
        torch, np.dtype(golden.np_dtype(stmt.haoda_type)).name))

  def read_input(name, box):
    index, dtype = dtypes[name]
    return synthetic(box, index, dtype, cpu).numpy()

  t0 = time.perf_counter()
  report = cone.check_windows(st, extent, read_input, read_output, windows)
  report['seconds'] = round(time.perf_counter() - t0, 3)
  return report


def device_reader(tensors, names, origin=0):
  """read_output over device arrays whose streamed axis starts at global
  slice ``origin``."""

  def read(name, box):
    tensor = tensors[names.index(name)]
    index = [slice(a, b) for a, b in reversed(box)]
    index[0] = slice(box[-1][0] - origin, box[-1][1] - origin)
    return tensor[tuple(index)].cpu().numpy()

  return read


def merge_parity(reports):
  total = {'windows': 0, 'cells': 0, 'bit_exact': True, 'first_mismatch': None}
  for r in reports:
    total['windows'] += r['windows']
    total['cells'] += r['cells']
    total['bit_exact'] = total['bit_exact'] and r['bit_exact']
    total['first_mismatch'] = total['first_mismatch'] or r['first_mismatch']
  return total


# ---------------------------------------------------------------------------
# one device-resident configuration on one GPU
# ---------------------------------------------------------------------------
def alu_ceiling(st, prog, extent):
  """Arithmetic ceiling of the pass kernel in Gcell-updates/s - the lower of
  the issue-slot and the FMA-pipe ceilings at 100 % utilisation, halo
  redundancy on this grid included - from the planner's instruction model
  (soda_b200/codegen/cuda/model.py)."""
  try:
    from soda_b200.codegen.cuda import model
    est = model.estimate_pass(st, prog.pass_info(0).time_block, None,
                              list(extent))
    return (min(est['issue_ceiling_gcells'], est['fma_ceiling_gcells']),
            est['instr_per_update'])
  except Exception:  # pylint: disable=broad-except
    return None, None


def run_device_config(key, device, stream, steps, warmup, peak, windows=8):
  """Times ``plan.run_device`` of a config with device-resident synthetic
  inputs and checks the measured output against the oracle."""
  import torch
  from soda_b200.codegen.cuda import launcher
  st, prog = config_program(key)
  extent = CONFIGS[key]['extent']
  dim = len(extent)
  shape = tuple(extent[::-1])
  ins = []
  for index, dt in enumerate(prog.input_dtypes):
    t = torch.empty(shape, dtype=getattr(torch, str(dt)), device=device)
    fill_synthetic(t, 0, index, extent[0])
    ins.append(t)
  outs = [torch.zeros(shape, dtype=getattr(torch, str(dt)), device=device)
          for dt in prog.output_dtypes]
  plane = extent[0] * extent[1] if dim == 3 else 0
  pitches = [(extent[0], plane)]
  plan = prog.create_plan(extent, launcher.make_opts(device=device.index,
                                                     stream=stream))

  def step():
    plan.run_device([t.data_ptr() for t in ins], pitches * len(ins),
                    [t.data_ptr() for t in outs], pitches * len(outs))

  for _ in range(warmup):
    step()
  torch.cuda.synchronize()
  before = prog.launch_count()
  start = torch.cuda.Event(enable_timing=True)
  end = torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(steps):
    step()
  end.record()
  torch.cuda.synchronize()
  launches = prog.launch_count() - before
  ms = start.elapsed_time(end) / steps
  cells = 1
  for e in extent:
    cells *= e
  passes = prog.num_passes
  gcells = cells * st.iterate / (ms * 1e-3) / 1e9
  achieved = cells * prog.bytes_per_cell_per_pass * passes / (ms * 1e-3) / 1e9
  time_block = prog.pass_info(0).time_block
  hbm_ceiling = peak / prog.bytes_per_cell_per_pass * st.iterate / passes
  alu, instr = alu_ceiling(st, prog, extent)
  parity = cone_parity(st, extent, device_reader(outs, prog.output_names),
                       windows, seed=3)
  result = {
      'config': key,
      'workload': '%s %s iterate %d%s' %
                  (CONFIGS[key]['program'] + (
                      ' --cuda-time-block %d' % CONFIGS[key]['time_block']
                      if CONFIGS[key].get('time_block') else '') + ''.join(
                          ' --cuda-%s' % k.replace('_', '-')
                          for k in (CONFIGS[key].get('options') or {})),
                   'x'.join(map(str, extent)),
                   st.iterate, ''.join(
                       ' --%s %s' % (k.replace('_', '-'), v)
                       for k, v in CONFIGS[key]['overrides'].items()
                       if k != 'iterate')),
      'value': gcells,
      'unit': 'Gcell-updates/s',
      'ms_per_step': ms,
      'time_block': time_block,
      'passes': passes,
      'steps': steps,
      'gpu_launches': launches,
      'roofline': {
          'bound': 'hbm',
          'achieved': achieved,
          'peak': peak,
          'unit': 'GB/s',
          'frac': achieved / peak,
          'hbm_ceiling_gcells': hbm_ceiling,
          'alu_ceiling_gcells': alu,
          'alu_model_instr_per_update': instr,
          'ceiling_gcells': min(hbm_ceiling, alu) if alu else hbm_ceiling,
          'frac_of_ceiling': gcells / (min(hbm_ceiling, alu)
                                       if alu else hbm_ceiling),
      },
      'parity': parity,
      'library': os.path.basename(prog.lib_path),
  }
  plan.close()
  return result, (st, prog, ins, outs)


# ---------------------------------------------------------------------------
# host <-> device copy peak of this box (the ceiling of e2e)
# ---------------------------------------------------------------------------
def measure_copy_peak(device, mib=256, reps=3, barrier=None):
  """GB/s of pinned host <-> device copies on this GPU: each direction alone
  and both at once (two streams).  With ``barrier`` (N > 1) all ranks copy at
  the same time, so the figures are what the box gives N GPUs together."""
  import torch
  n = (mib << 20) // 4
  h_a = torch.empty(n, dtype=torch.float32).pin_memory()
  h_b = torch.empty(n, dtype=torch.float32).pin_memory()
  h_a.fill_(1.0)
  d_a = torch.empty(n, dtype=torch.float32, device=device)
  d_b = torch.ones(n, dtype=torch.float32, device=device)
  s_in, s_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
  result = {}
  for mode in ('h2d', 'd2h', 'both'):
    best = None
    for rep in range(reps + 1):
      torch.cuda.synchronize()
      if barrier is not None:
        barrier()
      t0 = time.perf_counter()
      if mode in ('h2d', 'both'):
        with torch.cuda.stream(s_in):
          d_a.copy_(h_a, non_blocking=True)
      if mode in ('d2h', 'both'):
        with torch.cuda.stream(s_out):
          h_b.copy_(d_b, non_blocking=True)
      torch.cuda.synchronize()
      dt = time.perf_counter() - t0
      if rep > 0:
        best = dt if best is None else min(best, dt)
    nbytes = (mib << 20) * (2 if mode == 'both' else 1)
    result[mode + '_gbs'] = nbytes / best / 1e9
  return result


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(args, out):
  import numpy as np
  import torch
  import torch.distributed as dist
  from soda_b200.codegen.cuda import launcher, multi_gpu

  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device: there is no CPU path')
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    import datetime
    # a rank that fails must not leave the others waiting for ten minutes
    dist.init_process_group('nccl', device_id=device,
                            timeout=datetime.timedelta(seconds=180))

  st, prog = config_program(HEADLINE)
  stream = torch.cuda.current_stream().cuda_stream
  cells_per_gpu = WIDTH * HEIGHT
  n_passes = prog.num_passes
  time_block = prog.pass_info(0).time_block
  peak, peak_kind = measured_hbm_peak()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(x):
    if world == 1:
      return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  def sum_over_ranks(x):
    if world == 1:
      return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())

  def gather_parity(report):
    """Sums the ranks' parity reports on rank 0."""
    if world == 1:
      return report
    t = torch.tensor([report['windows'], report['cells'],
                      0 if report['bit_exact'] else 1],
                     dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    merged = dict(report)
    merged.update(windows=int(t[0]), cells=int(t[1]), bit_exact=int(t[2]) == 0,
                  ranks=world)
    return merged

  global_extent = (WIDTH, HEIGHT * world)
  if world == 1:
    d_in = torch.empty((HEIGHT, WIDTH), dtype=torch.float32, device=device)
    fill_synthetic(d_in, 0, 0, WIDTH)
    d_out = torch.zeros_like(d_in)
    plan = prog.create_plan((WIDTH, HEIGHT),
                            launcher.make_opts(device=local_rank,
                                               stream=stream))
    pitches = [(WIDTH, 0)]
    runner = None

    def step():
      plan.run_device([d_in.data_ptr()], pitches, [d_out.data_ptr()], pitches)
  else:
    runner = multi_gpu.SlabRunner(prog, global_extent, device,
                                  rank=rank, world=world,
                                  stream_handle=stream, transport=TRANSPORT)
    lo, hi = runner.own
    fill_synthetic(runner.inputs[0][lo:hi], runner.begin, 0, WIDTH)
    step = runner.run

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()

  for _ in range(args.warmup):
    step()
  barrier()
  launches_before = prog.launch_count()
  start = torch.cuda.Event(enable_timing=True)
  end = torch.cuda.Event(enable_timing=True)
  wall0 = time.perf_counter()
  start.record()
  for _ in range(args.steps):
    step()
  end.record()
  barrier()
  wall1 = time.perf_counter()
  launches = prog.launch_count() - launches_before
  ms_per_step = max_over_ranks(start.elapsed_time(end)) / args.steps
  value = cells_per_gpu * world * ITERATE / (ms_per_step * 1e-3) / 1e9

  # ---- parity of the measured output --------------------------------------
  parity = None
  if not args.no_parity:
    if world == 1:
      parity = cone_parity(st, global_extent,
                           device_reader([d_out], prog.output_names), 16)
    else:
      lo, hi = runner.own
      # two windows hug each slab boundary this rank has, the rest is random
      required = []
      if rank > 0:
        required += [(WIDTH // 3, runner.begin), (2 * WIDTH // 3, runner.begin)]
      if rank < world - 1:
        required += [(WIDTH // 3, runner.end - 64),
                     (2 * WIDTH // 3, runner.end - 64)]
      outputs = [runner.outputs[0][lo:hi]]
      parity = gather_parity(cone_parity(
          st, global_extent,
          device_reader(outputs, prog.output_names, origin=runner.begin),
          4, seed=rank, required=required, rows=(runner.begin, runner.end)))
      parity['slab_boundary_windows_per_rank'] = 'up to 4'

  # ---- end to end: host arrays through the public API (rank-local slab) ----
  e2e = None
  if args.no_e2e:
    pass
  elif world == 1:
    pinned_in = launcher.HostBuffer(prog, (HEIGHT, WIDTH), np.float32, local_rank)
    pinned_out = launcher.HostBuffer(prog, (HEIGHT, WIDTH), np.float32,
                                     local_rank)
    np_in, np_out = pinned_in.array, pinned_out.array
    torch.from_numpy(np_in).copy_(d_in)
    e2e_steps = max(1, min(args.steps, 5))
    plan.run_host({'t1': np_in}, {'t0': np_out})  # warm-up (allocations)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
      plan.run_host({'t1': np_in}, {'t0': np_out})
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    box = prog.valid_box(0, (WIDTH, HEIGHT))
    d2h = 4
    for lo_, hi_ in box:
      d2h *= max(0, hi_ - lo_)
    copy_peak = measure_copy_peak(device)
    e2e = {
        'value': cells_per_gpu * ITERATE / e2e_s / 1e9,
        'unit': 'Gcell-updates/s',
        'h2d_bytes_per_step': cells_per_gpu * 4,
        'd2h_bytes_per_step': d2h,
        'ms_per_step': e2e_s * 1e3,
        'host_memory': 'pinned (soda_cuda_host_alloc)',
        'api': 'soda_cuda_plan_run_host (chunked H2D / passes / D2H pipeline)',
        'chunks': 'automatic: 20 along the streamed dimension, each 8 % shorter '
                  'than the one before it',
        # the ceiling of this number: both copies at once on this box
        'pcie_peak_gbs': copy_peak['both_gbs'],
        'pcie_h2d_gbs': copy_peak['h2d_gbs'],
        'pcie_d2h_gbs': copy_peak['d2h_gbs'],
        'achieved_gbs': (cells_per_gpu * 4 + d2h) / e2e_s / 1e9,
        'frac': (cells_per_gpu * 4 + d2h) / e2e_s / 1e9 / copy_peak['both_gbs'],
        # the same run must agree with the device-resident one
        'checksum_matches_device_run': bool(
            np.array_equal(np_out[box[1][0]:box[1][1], box[0][0]:box[0][1]],
                           d_out[box[1][0]:box[1][1],
                                 box[0][0]:box[0][1]].cpu().numpy())),
    }
    if not args.no_parity:
      from oracle import cone
      e2e['parity'] = cone_parity(
          st, global_extent, cone.host_reader({'t0': np_out}), 8, seed=11)
    del np_in, np_out
    pinned_in.close()
    pinned_out.close()
  else:
    # N > 1: every rank's own slab in pinned host memory, through the slab's
    # chunked H2D / passes / D2H pipeline; the input ghosts (the total reach of
    # all passes) come from the neighbours' uploads over NVLink, once per step
    del runner
    torch.cuda.empty_cache()
    host_runner = multi_gpu.SlabRunner(prog, global_extent, device, rank=rank,
                                       world=world, stream_handle=stream,
                                       exchange_every=-1, transport=TRANSPORT)
    rows = host_runner.end - host_runner.begin
    pinned_in = launcher.HostBuffer(prog, (rows, WIDTH), np.float32, local_rank)
    pinned_out = launcher.HostBuffer(prog, (rows, WIDTH), np.float32,
                                     local_rank)
    np_in, np_out = pinned_in.array, pinned_out.array
    staging = torch.empty((rows, WIDTH), dtype=torch.float32, device=device)
    fill_synthetic(staging, host_runner.begin, 0, WIDTH)
    torch.from_numpy(np_in).copy_(staging)
    del staging
    e2e_steps = max(1, min(args.steps, 5))
    host_runner.run_host({'t1': np_in}, {'t0': np_out})  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
      host_runner.run_host({'t1': np_in}, {'t0': np_out})
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    copy_peak = measure_copy_peak(device, barrier=barrier)
    peak_all = sum_over_ranks(copy_peak['both_gbs'])
    nbytes = cells_per_gpu * 4 * world * 2
    e2e = {
        'value': cells_per_gpu * world * ITERATE / e2e_s / 1e9,
        'unit': 'Gcell-updates/s',
        'h2d_bytes_per_step': cells_per_gpu * 4 * world,
        'd2h_bytes_per_step': cells_per_gpu * 4 * world,
        'ms_per_step': e2e_s * 1e3,
        'host_memory': 'pinned (soda_cuda_host_alloc), one slab per rank',
        'api': 'soda_cuda_slab_run_host (per-rank chunked H2D / passes / D2H '
               'pipeline, one NVLink halo exchange per step)',
        'chunks': 'automatic: 20 per rank along the streamed dimension, each 8 % '
                  'shorter than the one before it',
        # all ranks copying both ways at the same time on this box
        'pcie_peak_gbs': peak_all,
        'achieved_gbs': nbytes / e2e_s / 1e9,
        'frac': nbytes / e2e_s / 1e9 / peak_all,
    }
    if not args.no_parity:
      from oracle import cone
      required = []
      if rank > 0:
        required.append((WIDTH // 2, host_runner.begin))
      if rank < world - 1:
        required.append((WIDTH // 2, host_runner.end - 64))

      def read_host(name, box, origin=host_runner.begin):
        del name
        return np_out[box[1][0] - origin:box[1][1] - origin,
                      box[0][0]:box[0][1]]

      e2e['parity'] = gather_parity(cone_parity(
          st, global_extent, read_host, 2, seed=50 + rank, required=required,
          rows=(host_runner.begin, host_runner.end)))
    del np_in, np_out
    pinned_in.close()
    pinned_out.close()
    host_runner.close()
    runner = None

  # ---- release the headline arrays, then the other BASELINE configs ---------
  if world == 1:
    plan.close()
    del d_in, d_out
  elif runner is not None:
    runner.close()
    del runner
  torch.cuda.empty_cache()

  other = []
  if not args.no_other and world == 1:
    for key in OTHER:
      try:
        result, keep = run_device_config(key, device, stream,
                                         steps=max(2, min(args.steps, 5)),
                                         warmup=max(3, min(args.warmup, 3)),
                                         peak=peak)
        del keep
      except Exception as e:  # pylint: disable=broad-except
        result = {'config': key, 'error': '%s: %s' % (type(e).__name__, e)}
      other.append(result)
      torch.cuda.empty_cache()

  c5 = None
  if not args.no_c5:
    try:
      c5 = run_c5_strong(args, device, stream, rank, world, peak, barrier,
                         max_over_ranks, gather_parity)
    except Exception as e:  # pylint: disable=broad-except
      c5 = {'error': '%s: %s' % (type(e).__name__, e)}

  if rank == 0:
    sampler.stop()
    ms_per_launch = ms_per_step / n_passes
    achieved = cells_per_gpu * prog.bytes_per_cell_per_pass / (
        ms_per_launch * 1e-3) / 1e9
    traffic = None
    traffic_path = os.path.join(ROOT, 'profiles',
                                'traffic_bytes_per_launch.json')
    if os.path.exists(traffic_path):
      # keyed by the library's hash-named file: a changed kernel has no entry
      try:
        with open(traffic_path) as fp:
          traffic = json.load(fp).get(os.path.basename(prog.lib_path))
      except Exception:  # pylint: disable=broad-except
        traffic = None
    line = {
        'metric': 'Gcell-updates/s',
        'value': value,
        'unit': 'Gcell-updates/s',
        'n_gpus': world,
        'steps': args.steps,
        'warmup': args.warmup,
        'ms_per_step': ms_per_step,
        'higher_is_better': True,
        'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f32',
        'data': 'synthetic (integer hash of the cell coordinates -> U[0,1))',
        'config': workload_config(world, time_block, n_passes),
        'roofline': {
            'bound': 'hbm',
            'kernel': 'soda_stream2d_kernel<jacobi2d, time_block=%d>' %
                      time_block,
            'library': os.path.basename(prog.lib_path),
            'achieved': achieved,
            'peak': peak,
            'peak_source': peak_kind,
            'unit': 'GB/s',
            'frac': achieved / peak,
            'traffic': traffic,
            'algorithmic_bytes_per_launch': cells_per_gpu *
                                            prog.bytes_per_cell_per_pass,
            'ms_per_launch': ms_per_launch,
            'launches_per_step': n_passes,
            'time_block': time_block,
            'gcell_updates_per_s_vs_time_block_1_ceiling':
                value / world / (peak / prog.bytes_per_cell_per_pass),
        },
        'parity': parity,
        'e2e': e2e,
        'gpu_launches': launches,
        'clocks': sampler.summary(wall0, wall1),
        'wall_ms_per_step': (wall1 - wall0) / args.steps * 1e3,
        'other_configs': other,
        'c5_strong': c5,
    }
    if world == 1 and not args.no_cpu_baseline:
      line['cpu_baseline'] = cpu_baseline()
      try:
        line['cpu_baseline_dataflow'] = cpu_baseline_dataflow()
      except Exception as e:  # pylint: disable=broad-except
        line['cpu_baseline_dataflow'] = {
            'error': '%s: %s' % (type(e).__name__, e)}
    out.emit(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def run_c5_strong(args, device, stream, rank, world, peak, barrier,
                  max_over_ranks, gather_parity):
  """BASELINE config C5: one 65536 x 65536 grid, iterate 256, split over the
  ranks of this run along the streamed dimension (strong scaling)."""
  import torch
  from soda_b200.codegen.cuda import launcher, multi_gpu
  key = 'C5_jacobi2d'
  st, prog = config_program(key)
  width, height = CONFIGS[key]['extent']
  iterate = st.iterate
  steps = 2
  exchange = None
  if world == 1:
    d_in = torch.empty((height, width), dtype=torch.float32, device=device)
    fill_synthetic(d_in, 0, 0, width)
    d_out = torch.zeros_like(d_in)
    plan = prog.create_plan((width, height),
                            launcher.make_opts(device=device.index,
                                               stream=stream))

    def step():
      plan.run_device([d_in.data_ptr()], [(width, 0)], [d_out.data_ptr()],
                      [(width, 0)])
    outputs, origin, rows = [d_out], 0, (0, height)
    required = []
    groups = 1
  else:
    runner = multi_gpu.SlabRunner(prog, (width, height), device, rank=rank,
                                  world=world, stream_handle=stream,
                                  transport=TRANSPORT)
    lo, hi = runner.own
    fill_synthetic(runner.inputs[0][lo:hi], runner.begin, 0, width)
    step = runner.run
    outputs, origin = [runner.outputs[0][lo:hi]], runner.begin
    rows = (runner.begin, runner.end)
    required = []
    if rank > 0:
      required.append((width // 2, runner.begin))
    if rank < world - 1:
      required.append((width // 2, runner.end - 64))
    groups = len(runner.groups)
  step()  # warm-up (segment measurement, scratch allocation)
  step()
  barrier()
  before = prog.launch_count()
  start = torch.cuda.Event(enable_timing=True)
  end = torch.cuda.Event(enable_timing=True)
  start.record()
  for _ in range(steps):
    step()
  end.record()
  barrier()
  launches = prog.launch_count() - before
  ms = max_over_ranks(start.elapsed_time(end)) / steps
  if world > 1:
    # the cost of one halo exchange of a full group's depth, on its own
    depth_lo = max(sum(runner.pass_reach[i][0] for i in g)
                   for g in runner.groups)
    depth_hi = max(sum(runner.pass_reach[i][1] for i in g)
                   for g in runner.groups)
    runner.exchange(runner.inputs, depth_lo, depth_hi)
    barrier()
    start.record()
    for _ in range(5):
      runner.exchange(runner.inputs, depth_lo, depth_hi)
    end.record()
    barrier()
    exchange = {
        'groups_per_step': groups,
        'passes_per_group': [len(g) for g in runner.groups][:3],
        'ghost_rows': [depth_lo, depth_hi],
        'bytes_per_direction': depth_hi * width * 4,
        'ms_alone': max_over_ranks(start.elapsed_time(end)) / 5,
        'overlapped_with': 'the interior launch of the group\'s last pass',
    }
  parity = None
  if not args.no_parity:
    parity = gather_parity(cone_parity(
        st, (width, height),
        device_reader(outputs, prog.output_names, origin=origin),
        2 if world > 1 else 4, seed=100 + rank, required=required, rows=rows))
  cells = width * height
  passes = prog.num_passes
  achieved = cells * prog.bytes_per_cell_per_pass * passes / (ms * 1e-3) / 1e9
  return {
      'workload': 'jacobi2d fp32 %dx%d iterate %d' % (width, height, iterate),
      'scaling': 'strong',
      'n_gpus': world,
      'slab_rows_per_gpu': height // world,
      'value': cells * iterate / (ms * 1e-3) / 1e9,
      'unit': 'Gcell-updates/s',
      'ms_per_step': ms,
      'steps': steps,
      'time_block': prog.pass_info(0).time_block,
      'passes': passes,
      'gpu_launches': launches,
      'roofline_frac_per_gpu': achieved / world / peak,
      'exchange': exchange,
      'parity': parity,
  }


class JsonOnlyStdout:
  """Sends everything libraries print to stdout (e.g. NCCL's version banner)
  to stderr, so that stdout carries exactly the one JSON line."""

  def __enter__(self):
    sys.stdout.flush()
    self.saved = os.dup(1)
    os.dup2(2, 1)
    return self

  def emit(self, line):
    sys.stdout.flush()
    os.write(self.saved, (line + '\n').encode())

  def __exit__(self, *exc):
    sys.stdout.flush()
    os.dup2(self.saved, 1)
    os.close(self.saved)


def main():
  parser = argparse.ArgumentParser()
  parser.add_argument('--gpus', type=int, default=1)
  parser.add_argument('--steps', type=int, default=20)
  parser.add_argument('--warmup', type=int, default=3)
  parser.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  parser.add_argument('--no-cpu-baseline', action='store_true')
  parser.add_argument('--no-e2e', action='store_true',
                      help='profiling runs: only the device-resident steps')
  parser.add_argument('--no-parity', action='store_true')
  parser.add_argument('--no-other', action='store_true',
                      help='skip the other BASELINE configs (C1, C3, C4)')
  parser.add_argument('--no-c5', action='store_true',
                      help='skip BASELINE config C5 (65536^2, iterate 256)')
  parser.add_argument('--headline-only', action='store_true',
                      help='= --no-e2e --no-parity --no-other --no-c5 '
                      '--no-cpu-baseline (profiling under ncu)')
  args = parser.parse_args()
  if args.headline_only:
    args.no_e2e = args.no_parity = args.no_other = args.no_c5 = True
    args.no_cpu_baseline = True
  with JsonOnlyStdout() as out:
    if args.impl == 'reference':
      run_reference(args, out)
    else:
      run_ours(args, out)


if __name__ == '__main__':
  main()
