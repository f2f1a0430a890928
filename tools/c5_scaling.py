#!/usr/bin/env python3
"""BASELINE config C5: jacobi2d fp32 on a 65536 x 65536 grid, iterate 256.

  python tools/c5_scaling.py                     # 1 GPU
  torchrun --nproc-per-node N tools/c5_scaling.py  # strong scaling over N GPUs

Checks ("Large-grid parity"):
  * N > 1: every rank also runs the whole grid alone and compares its slab of
    the N-GPU result with the 1-GPU result bit for bit;
  * rank 0: 8 random 64 x 64 windows of the 1-GPU result against the CPU oracle
    evaluated on each window's dependency cone (64 + 2*iterate cells wide).
Prints one JSON line per rank-0 run.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen import cuda as cb  # noqa: E402
from soda_b200.codegen.cuda import launcher, multi_gpu  # noqa: E402

W = H = int(os.environ.get('C5_SIZE', '65536'))
ITERATE = int(os.environ.get('C5_ITERATE', '256'))
TB = int(os.environ.get('C5_TIME_BLOCK', '6'))


def main():
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local_rank)
  device = torch.device('cuda', local_rank)
  if world > 1:
    dist.init_process_group('nccl', device_id=device)
  with open(os.path.join(ROOT, 'tests', 'src', 'jacobi2d.soda')) as fp:
    st = sodac.compile_source(fp.read(), iterate=ITERATE)
  prog = cb.compile_stencil(st, time_block=TB)
  stream = torch.cuda.current_stream().cuda_stream

  # the same global input on every rank (seeded generator, identical devices)
  gen = torch.Generator(device=device)
  gen.manual_seed(5)
  full_in = torch.rand((H, W), dtype=torch.float32, device=device,
                       generator=gen)
  full_out = torch.zeros_like(full_in)
  plan = prog.create_plan((W, H), launcher.make_opts(device=local_rank,
                                                     stream=stream))
  run1 = lambda: plan.run_device([full_in.data_ptr()], [(W, 0)],
                                 [full_out.data_ptr()], [(W, 0)])
  run1()
  torch.cuda.synchronize()
  start, end = torch.cuda.Event(enable_timing=True), \
      torch.cuda.Event(enable_timing=True)
  start.record()
  run1()
  end.record()
  torch.cuda.synchronize()
  ms_single = start.elapsed_time(end)
  result = {'grid': [W, H], 'iterate': ITERATE, 'time_block': TB,
            'n_gpus': world, 'ms_1gpu': ms_single,
            'gcell_per_s_1gpu': W * H * ITERATE / ms_single / 1e6}

  if world > 1:
    plan.close()
    runner = multi_gpu.SlabRunner(prog, (W, H), device, rank=rank, world=world,
                                  stream_handle=stream)
    lo, hi = runner.own
    runner.view(runner.inputs[0])[lo:hi].copy_(
        full_in[runner.begin:runner.end])
    runner.view(runner.outputs[0]).zero_()
    runner.run()
    dist.barrier()
    torch.cuda.synchronize()
    start.record()
    runner.run()
    end.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([start.elapsed_time(end)], dtype=torch.float64,
                     device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_multi = float(t.item())
    mine = runner.view(runner.outputs[0])[lo:hi]
    same = torch.equal(mine.view(torch.int32),
                       full_out[runner.begin:runner.end].view(torch.int32))
    flag = torch.tensor([1 if same else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    result.update({'ms_ngpu': ms_multi,
                   'gcell_per_s_ngpu': W * H * ITERATE / ms_multi / 1e6,
                   'strong_scaling_efficiency': ms_single / ms_multi / world,
                   'bitwise_equal_to_1gpu': bool(flag.item())})

  if rank == 0:
    from oracle import emit_cpp
    oracle = emit_cpp.Oracle(st)
    rng = np.random.default_rng(7)
    r = ITERATE
    ok = True
    windows = int(os.environ.get('C5_WINDOWS', '8'))
    t0 = time.perf_counter()
    for _ in range(windows):
      y0 = int(rng.integers(r, H - 64 - r))
      x0 = int(rng.integers(r, W - 64 - r))
      sub = full_in[y0 - r:y0 + 64 + r, x0 - r:x0 + 64 + r].cpu().numpy()
      want = oracle.run({'t1': np.ascontiguousarray(sub)})['t0'][r:-r, r:-r]
      got = full_out[y0:y0 + 64, x0:x0 + 64].cpu().numpy()
      ok = ok and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    result['oracle_windows_bit_exact'] = bool(ok)
    result['oracle_windows'] = windows
    result['oracle_seconds'] = time.perf_counter() - t0
    print(json.dumps(result), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()
