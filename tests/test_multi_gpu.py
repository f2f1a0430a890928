"""Slab decomposition + halo exchange, world_size 2 and 3, on the CPU: gloo for
the exchange, the emulated program library (tests/emu) for the passes.  The
N-rank result must be bit-identical to the single-rank oracle on the whole
global grid ("N-GPU result == 1-GPU result bit-for-bit")."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import common


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


def _worker(rank, world, port, name, overrides, time_block, extent, lib, seed,
            result_dir, exchange_every=None, expect_groups=None):
  sys.path.insert(0, common.ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from soda_b200.codegen.cuda import launcher, multi_gpu
    st = common.stencil(name, **overrides)
    prog = launcher.CudaProgram(lib)
    runner = multi_gpu.SlabRunner(prog, extent, torch.device('cpu'), rank=rank,
                                  world=world, exchange_every=exchange_every)
    if expect_groups is not None:
      assert [len(g) for g in runner.groups] == expect_groups, runner.groups
    inputs = common.make_inputs(st, extent, seed=seed)
    lo, hi = runner.own
    for tensor, iname in zip(runner.inputs, st.input_names):
      full = torch.from_numpy(inputs[iname])
      runner.view(tensor)[lo:hi].copy_(full[runner.begin:runner.end])
    for tensor in runner.outputs:
      tensor.fill_(77)
    runner.run()
    assert runner.launches >= prog.num_passes
    for tensor, oname in zip(runner.outputs, st.output_names):
      np.save(os.path.join(result_dir, '%s_%d.npy' % (oname, rank)),
              runner.view(tensor)[lo:hi].numpy())
  finally:
    dist.destroy_process_group()


def run_case(tmp_path, name, extent, world, time_block=None, options=None,
             seed=3, exchange_every=None, expect_groups=None, **overrides):
  from tests.emu import build_emu
  st = common.stencil(name, **overrides)
  lib = build_emu.build_emu_library(st, time_block=time_block, options=options)
  port = _free_port()
  mp.spawn(_worker,
           args=(world, port, name, overrides, time_block, extent, lib, seed,
                 str(tmp_path), exchange_every, expect_groups),
           nprocs=world, join=True)
  inputs = common.make_inputs(st, extent, seed=seed)
  want = common.oracle_outputs(st, inputs)
  for oname in st.output_names:
    parts = [np.load(os.path.join(str(tmp_path), '%s_%d.npy' % (oname, r)))
             for r in range(world)]
    got = np.concatenate(parts, axis=0)
    common.assert_matches_oracle(st, extent, {oname: got}, want, sentinel=77)


def test_split_slices():
  from soda_b200.codegen.cuda import multi_gpu
  assert multi_gpu.split_slices(10, 3) == [(0, 4), (4, 7), (7, 10)]
  assert multi_gpu.split_slices(8, 2) == [(0, 4), (4, 8)]


def test_jacobi2d_two_ranks(tmp_path):
  run_case(tmp_path, 'jacobi2d', (140, 61), 2, time_block=2, iterate=5)


def test_jacobi2d_three_ranks_time_block_4(tmp_path):
  run_case(tmp_path, 'jacobi2d', (70, 90), 3, time_block=4, iterate=9)


def test_one_sided_window_two_ranks(tmp_path):
  run_case(tmp_path, 'blur', (300, 40), 2, time_block=2, iterate=2)


def test_multi_input_dag_two_ranks(tmp_path):
  run_case(tmp_path, 'denoise2d', (64, 30), 2)


def test_heat3d_two_ranks(tmp_path):
  run_case(tmp_path, 'heat3d', (40, 12, 20), 2, time_block=2, iterate=4,
           options={'rows': 8})


@pytest.mark.parametrize('every,groups', [(2, [2, 2, 1]), (5, [5]), (1, [1] * 5)])
def test_exchange_groups_three_ranks(tmp_path, every, groups):
  """Several passes between two halo exchanges: every pass of a group also
  computes the ghost slices the rest of the group still reads; results stay
  bit-identical to the single-rank oracle."""
  run_case(tmp_path, 'jacobi2d', (70, 90), 3, time_block=2, iterate=9,
           exchange_every=every, expect_groups=groups)


def test_exchange_groups_one_sided_and_3d(tmp_path):
  run_case(tmp_path, 'blur', (300, 60), 2, time_block=1, iterate=3,
           exchange_every=3, expect_groups=[3])
  run_case(tmp_path, 'heat3d', (40, 12, 30), 2, time_block=1, iterate=4,
           options={'rows': 8}, exchange_every=2, expect_groups=[2, 2])


def test_default_groups_keep_the_ghost_small():
  from soda_b200.codegen.cuda import multi_gpu

  class FakeInfo:
    def __init__(self):
      self.reach_lo = [0, -6, 0]
      self.reach_hi = [0, 6, 0]

  class FakeProgram:
    dim = 2
    num_passes = 11
    input_dtypes = output_dtypes = []
    def pass_info(self, index):
      return FakeInfo()

  runner = multi_gpu.SlabRunner(FakeProgram(), (16384, 8 * 16384),
                                torch.device('cpu'), rank=3, world=8)
  # 66 ghost rows per side of a 16384-row slab: one exchange per step
  assert [len(g) for g in runner.groups] == [11]
  assert (runner.reach_lo, runner.reach_hi) == (66, 66)
  thin = multi_gpu.SlabRunner(FakeProgram(), (16384, 8 * 512),
                              torch.device('cpu'), rank=0, world=8)
  assert [len(g) for g in thin.groups] == [2, 2, 2, 2, 2, 1]  # 12 of 512 rows
