#!/usr/bin/env python3
"""Benchmark harness: Gcell-updates/s of the SODA CUDA backend on B200.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): jacobi2d, fp32 5-point, 16384 x 16384,
iterate 64, synthetic U[0,1) input.  One "step" = all 64 iterations over the
grid.  With N > 1 (launched by torchrun, one rank per GPU) every rank owns a
16384 x 16384 slab of a 16384 x (16384 N) grid (weak scaling) and swaps halo
rows with its neighbours before every pass.

Prints ONE JSON line (see the driver contract) carrying, besides the metric:
  roofline     - per-pass HBM roofline of the dominant kernel, from CUDA events
  cpu_baseline - the g++-compiled restatement of the reference's loops (the
                 oracle, kind "port") timed on this box's host cores
  e2e          - the same metric through the public host-array API, with the
                 host->device and device->host copies inside the timed region
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

WIDTH = 16384
HEIGHT = 16384
ITERATE = 64
TIME_BLOCK = int(os.environ.get('SODA_BENCH_TIME_BLOCK', '6'))
PROGRAM = 'jacobi2d'
BYTES_PER_CELL_PER_PASS = 8  # one fp32 read + one fp32 write
FALLBACK_HBM_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md


def workload_config(n_gpus):
  return {
      'workload': 'jacobi2d fp32 5-point %dx%d iterate %d' %
                  (WIDTH, HEIGHT * n_gpus, ITERATE),
      'program': 'tests/src/jacobi2d.soda --iterate %d' % ITERATE,
      'grid_per_gpu': [WIDTH, HEIGHT],
      'time_block': TIME_BLOCK,
      'passes': -(-ITERATE // TIME_BLOCK),
      'parallelism': 'slab%d' % n_gpus if n_gpus > 1 else 'single',
      'l2': 'inputs (1 GiB per array per GPU) are larger than the 126 MB L2',
  }


def measured_hbm_peak():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  try:
    with open(path) as fp:
      return float(json.load(fp)['hbm_gbs']), 'measured'
  except Exception:  # pylint: disable=broad-except
    return FALLBACK_HBM_GBS, 'fallback'


def stencil(iterate=ITERATE):
  from soda_b200 import sodac
  with open(os.path.join(ROOT, 'tests', 'src', PROGRAM + '.soda')) as fp:
    return sodac.compile_source(fp.read(), iterate=iterate)


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
# CPU baseline (the oracle's g++ build; the only place bench.py runs oracle/)
# ---------------------------------------------------------------------------
_CPU_STATE = {}


def cpu_baseline(sample_iterate=4, repeats=2):
  import numpy as np
  from oracle import emit_cpp
  if sample_iterate not in _CPU_STATE:
    st = stencil(iterate=sample_iterate)
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
# our arm
# ---------------------------------------------------------------------------
def run_ours(args, out):
  import numpy as np
  import torch
  import torch.distributed as dist
  from soda_b200.codegen import cuda as cuda_backend
  from soda_b200.codegen.cuda import launcher, multi_gpu

  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device: there is no CPU path')
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=device)

  prog = cuda_backend.compile_stencil(stencil(), time_block=TIME_BLOCK)
  stream = torch.cuda.current_stream().cuda_stream
  cells_per_gpu = WIDTH * HEIGHT
  n_passes = prog.num_passes

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  gen = torch.Generator(device=device)
  gen.manual_seed(1 + rank)

  if world == 1:
    d_in = torch.rand((HEIGHT, WIDTH), dtype=torch.float32, device=device,
                      generator=gen)
    d_out = torch.zeros_like(d_in)
    plan = prog.create_plan((WIDTH, HEIGHT),
                            launcher.make_opts(device=local_rank,
                                               stream=stream))
    pitches = [(WIDTH, 0)]

    def step():
      plan.run_device([d_in.data_ptr()], pitches, [d_out.data_ptr()], pitches)
  else:
    runner = multi_gpu.SlabRunner(prog, (WIDTH, HEIGHT * world), device,
                                  rank=rank, world=world,
                                  stream_handle=stream)
    lo, hi = runner.own
    runner.view(runner.inputs[0])[lo:hi].copy_(
        torch.rand((hi - lo, WIDTH), dtype=torch.float32, device=device,
                   generator=gen))
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
  elapsed_ms = start.elapsed_time(end)
  if world > 1:
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
  ms_per_step = elapsed_ms / args.steps
  value = cells_per_gpu * world * ITERATE / (ms_per_step * 1e-3) / 1e9

  # ---- end to end: host arrays through the public API (rank-local slab) ----
  e2e = None
  if args.no_e2e:
    pass
  elif world == 1:
    h_in = torch.empty((HEIGHT, WIDTH), dtype=torch.float32).pin_memory()
    h_in.copy_(d_in)
    h_out = torch.zeros((HEIGHT, WIDTH), dtype=torch.float32).pin_memory()
    np_in, np_out = h_in.numpy(), h_out.numpy()
    e2e_steps = max(1, min(args.steps, 3))
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
    e2e = {
        'value': cells_per_gpu * ITERATE / e2e_s / 1e9,
        'unit': 'Gcell-updates/s',
        'h2d_bytes_per_step': cells_per_gpu * 4,
        'd2h_bytes_per_step': d2h,
        'ms_per_step': e2e_s * 1e3,
        'host_memory': 'pinned',
        # the same run must agree with the device-resident one
        'checksum_matches_device_run': bool(
            np.array_equal(np_out[box[1][0]:box[1][1], box[0][0]:box[0][1]],
                           d_out[box[1][0]:box[1][1],
                                 box[0][0]:box[0][1]].cpu().numpy())),
    }
  else:
    # N > 1: every rank stages its own slab through pinned memory
    lo, hi = runner.own
    own = runner.view(runner.inputs[0])[lo:hi]
    h_in = torch.empty(own.shape, dtype=torch.float32).pin_memory()
    h_in.copy_(own)
    h_out = torch.empty(own.shape, dtype=torch.float32).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
      own.copy_(h_in, non_blocking=True)
      runner.run()
      h_out.copy_(runner.view(runner.outputs[0])[lo:hi], non_blocking=True)
      torch.cuda.synchronize()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e = {
        'value': cells_per_gpu * world * ITERATE / e2e_s / 1e9,
        'unit': 'Gcell-updates/s',
        'h2d_bytes_per_step': cells_per_gpu * 4 * world,
        'd2h_bytes_per_step': cells_per_gpu * 4 * world,
        'ms_per_step': e2e_s * 1e3,
        'host_memory': 'pinned',
    }

  if rank == 0:
    sampler.stop()
    peak, peak_kind = measured_hbm_peak()
    ms_per_launch = ms_per_step / n_passes
    achieved = cells_per_gpu * BYTES_PER_CELL_PER_PASS / (ms_per_launch *
                                                          1e-3) / 1e9
    traffic = None
    traffic_path = os.path.join(ROOT, 'profiles', 'traffic_bytes_per_launch.json')
    if os.path.exists(traffic_path):
      try:
        with open(traffic_path) as fp:
          traffic = json.load(fp).get('jacobi2d_16384_tb%d' % TIME_BLOCK)
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
        'data': 'synthetic',
        'config': workload_config(world),
        'roofline': {
            'bound': 'hbm',
            'kernel': 'soda_stream2d_kernel<jacobi2d, time_block=%d>' %
                      TIME_BLOCK,
            'achieved': achieved,
            'peak': peak,
            'peak_source': peak_kind,
            'unit': 'GB/s',
            'frac': achieved / peak,
            'traffic': traffic,
            'algorithmic_bytes_per_launch': cells_per_gpu *
                                            BYTES_PER_CELL_PER_PASS,
            'ms_per_launch': ms_per_launch,
            'launches_per_step': n_passes,
            'time_block': TIME_BLOCK,
            'gcell_updates_per_s_vs_time_block_1_ceiling':
                value / world / (peak / BYTES_PER_CELL_PER_PASS),
        },
        'e2e': e2e,
        'gpu_launches': launches,
        'clocks': sampler.summary(wall0, wall1),
        'wall_ms_per_step': (wall1 - wall0) / args.steps * 1e3,
    }
    if world == 1 and not args.no_cpu_baseline:
      line['cpu_baseline'] = cpu_baseline()
    out.emit(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


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
  args = parser.parse_args()
  with JsonOnlyStdout() as out:
    if args.impl == 'reference':
      run_reference(args, out)
    else:
      run_ours(args, out)


if __name__ == '__main__':
  main()
