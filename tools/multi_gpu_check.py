#!/usr/bin/env python3
"""N-GPU slab run == 1-GPU run, bit for bit, for programs other than the bench's:
a 3-D stencil with temporal blocking, a multi-input DAG, a one-sided uint16
window and a program with exchange groups of several passes.

  torchrun --nproc-per-node N tools/multi_gpu_check.py
  torchrun --nproc-per-node N tools/multi_gpu_check.py random LO HI [hard]
      the same check on the seeded random programs of tests/random_programs.py
      (asymmetric windows along the slab dimension, off-centre stores, mixed
      types), grids stretched along the last dimension, 1-3 passes per exchange

Every rank runs the whole grid alone (plan.run_device) and its slab of the
N-rank run (SlabRunner) and compares the two; rank 0 prints one JSON line per
case.  SURVEY section 8(d): "N-GPU result == 1-GPU result bit-for-bit".
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen import cuda as cb  # noqa: E402
from soda_b200.codegen.cuda import launcher, multi_gpu  # noqa: E402

CASES = [
    # program, overrides, time block, extent, exchange_every
    ('heat3d', {'iterate': 8}, 2, (256, 256, 512), None),
    ('heat3d', {'iterate': 8}, 2, (256, 256, 512), 1),
    ('jacobi3d', {'iterate': 6}, 1, (512, 512, 256), None),
    ('denoise2d', {}, None, (4096, 4096), None),
    ('blur', {'iterate': 4}, 2, (16000, 4096), None),
    ('jacobi2d', {'iterate': 64}, 6, (16384, 8192), 4),
    ('seidel2d', {'iterate': 12}, 4, (8192, 8192), None),
]


def random_cases(lo, hi, hard):
  from tests import random_programs
  cases = []
  for seed in range(lo, hi):
    text, extent, kwargs = random_programs.program(seed, hard=hard)
    # dense device arrays: TMA wants pitches that are multiples of 16 bytes
    extent = ((extent[0] + 15) // 16 * 16,) + tuple(extent[1:-1]) + (
        extent[-1] * 6,)
    every = (None, 1, 2, 3)[seed % 4]
    cases.append(('rnd%d%s' % (seed, 'h' if hard else ''), text,
                  kwargs.get('time_block'), kwargs.get('options'), extent,
                  every))
  return cases


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  dist.init_process_group('nccl', device_id=device)
  stream = torch.cuda.current_stream().cuda_stream
  cases = []
  for name, overrides, tb, extent, every in CASES:
    with open(os.path.join(ROOT, 'tests', 'src', name + '.soda')) as fp:
      cases.append((name, fp.read(), tb, None, extent, every, overrides))
  if len(sys.argv) > 3 and sys.argv[1] == 'random':
    cases = [c + ({},) for c in random_cases(int(sys.argv[2]), int(sys.argv[3]),
                                             'hard' in sys.argv[4:])]
  failures = 0
  for name, text, tb, options, extent, every, overrides in cases:
    from soda_b200 import util
    st = sodac.compile_source(text, **overrides)
    try:
      prog = cb.compile_stencil(st, time_block=tb, options=options)
    except util.SemanticError as e:
      if rank == 0:
        print(json.dumps(dict(program=name, skipped='plan: %s' % e)), flush=True)
      continue
    shape = tuple(extent[::-1])
    gen = torch.Generator(device=device)
    gen.manual_seed(11)
    full_in, full_out = [], []
    for dt in prog.input_dtypes:
      tdt = getattr(torch, str(dt))
      if dt.kind == 'f':
        full_in.append(torch.rand(shape, dtype=tdt, device=device,
                                  generator=gen))
      else:
        top = 60000 if name == 'blur' else 1000
        full_in.append(torch.randint(0, top, shape, device=device,
                                     generator=gen).to(torch.int32).to(tdt))
    for dt in prog.output_dtypes:
      full_out.append(torch.zeros(shape, dtype=getattr(torch, str(dt)),
                                  device=device))
    plane = extent[0] * extent[1] if len(extent) == 3 else 0
    pitches = [(extent[0], plane)]
    plan = prog.create_plan(extent, launcher.make_opts(device=local_rank,
                                                       stream=stream))
    plan.run_device([t.data_ptr() for t in full_in], pitches * len(full_in),
                    [t.data_ptr() for t in full_out], pitches * len(full_out))
    torch.cuda.synchronize()
    plan.close()

    try:
      runner = multi_gpu.SlabRunner(prog, extent, device, rank=rank,
                                    world=world, stream_handle=stream,
                                    exchange_every=every)
    except ValueError as e:  # slab thinner than the halo: a documented limit
      if rank == 0:
        print(json.dumps(dict(program=name, skipped=str(e))), flush=True)
      del full_in, full_out
      continue
    lo, hi = runner.own
    for local, full in zip(runner.inputs, full_in):
      runner.view(local)[lo:hi].copy_(full[runner.begin:runner.end])
    runner.run()
    torch.cuda.synchronize()
    same = True
    for o, (local, full) in enumerate(zip(runner.outputs, full_out)):
      box = prog.valid_box(o, extent)
      index = tuple(slice(b[0], b[1]) for b in reversed(box[:-1]))
      g_lo = max(runner.begin, box[-1][0])
      g_hi = min(runner.end, box[-1][1])
      if g_hi <= g_lo:
        continue
      mine = runner.view(local)[lo + (g_lo - runner.begin):
                                lo + (g_hi - runner.begin)][(slice(None),) + index]
      ref = full[g_lo:g_hi][(slice(None),) + index]
      same = same and torch.equal(mine.contiguous().view(torch.uint8),
                                  ref.contiguous().view(torch.uint8))
    flag = torch.tensor([1 if same else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    failures += 0 if flag.item() else 1
    if rank == 0:
      print(json.dumps(dict(program=name, overrides=overrides, time_block=tb,
                            options=options, exchange_every=every,
                            extent=extent, n_gpus=world,
                            exchange_groups=[len(g) for g in runner.groups],
                            bitwise_equal_to_1gpu=bool(flag.item()))),
            flush=True)
    del full_in, full_out, runner
    torch.cuda.empty_cache()
  if rank == 0:
    print(json.dumps(dict(cases=len(cases), not_equal=failures)), flush=True)
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
