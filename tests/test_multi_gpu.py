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
            result_dir, exchange_every=None, expect_groups=None,
            host_chunks=None, edge_chunks=None):
  sys.path.insert(0, common.ROOT)
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from soda_b200.codegen.cuda import launcher, multi_gpu
    st = common.stencil(name, **overrides)
    prog = launcher.CudaProgram(lib)
    runner = multi_gpu.SlabRunner(prog, extent, torch.device('cpu'), rank=rank,
                                  world=world, exchange_every=exchange_every,
                                  host_chunks=host_chunks or 0,
                                  edge_chunks=edge_chunks)
    if expect_groups is not None:
      assert [len(g) for g in runner.groups] == expect_groups, runner.groups
    inputs = common.make_inputs(st, extent, seed=seed)
    lo, hi = runner.own
    if host_chunks:
      # host arrays of this rank's own slices through the chunked pipeline
      own_in = {n: np.ascontiguousarray(inputs[n][runner.begin:runner.end])
                for n in st.input_names}
      own_out = {n: np.full((runner.end - runner.begin,) + extent[-2::-1], 77,
                            dtype=d)
                 for n, d in zip(prog.output_names, prog.output_dtypes)}
      runner.run_host(own_in, own_out)
      for oname in st.output_names:
        np.save(os.path.join(result_dir, '%s_%d.npy' % (oname, rank)),
                own_out[oname])
      return
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
             seed=3, exchange_every=None, expect_groups=None, host_chunks=None,
             edge_chunks=None, **overrides):
  from tests.emu import build_emu
  st = common.stencil(name, **overrides)
  lib = build_emu.build_emu_library(st, time_block=time_block, options=options)
  port = _free_port()
  mp.spawn(_worker,
           args=(world, port, name, overrides, time_block, extent, lib, seed,
                 str(tmp_path), exchange_every, expect_groups, host_chunks,
                 edge_chunks),
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
  """The library's grouping rule on the bench shapes, asked for as a dry run
  (nothing is allocated): 11 passes of reach 6 on 16384-row slabs are one
  group, on 512-row slabs groups of two."""
  from soda_b200.codegen.cuda import launcher, multi_gpu
  from tests.emu import build_emu
  st = common.stencil('jacobi2d', iterate=64)
  prog = launcher.CudaProgram(build_emu.build_emu_library(st, time_block=6))
  assert prog.num_passes == 11
  runner = multi_gpu.SlabRunner(prog, (16384, 8 * 16384), torch.device('cpu'),
                                rank=3, world=8, dry_run=True)
  # 64 ghost rows per side of a 16384-row slab: one exchange per step
  assert [len(g) for g in runner.groups] == [11]
  assert (runner.reach_lo, runner.reach_hi) == (64, 64)
  assert (runner.begin, runner.end) == (3 * 16384, 4 * 16384)
  assert runner.own == (64, 64 + 16384)
  thin = multi_gpu.SlabRunner(prog, (16384, 8 * 512), torch.device('cpu'),
                              rank=0, world=8, dry_run=True)
  assert [len(g) for g in thin.groups] == [2, 2, 2, 2, 2, 1]  # 12 of 512 rows
  assert thin.own == (0, 512) and thin.local_extent == (16384, 512 + 12)
  everything = multi_gpu.SlabRunner(prog, (16384, 8 * 512), torch.device('cpu'),
                                    rank=1, world=8, exchange_every=-1,
                                    dry_run=True)
  assert [len(g) for g in everything.groups] == [11]


@pytest.mark.parametrize('name,extent,world,kwargs', [
    ('jacobi2d', (70, 120), 3, dict(time_block=2, iterate=5, host_chunks=4)),
    ('jacobi2d', (70, 60), 2, dict(time_block=2, iterate=5, host_chunks=1)),
    # chunks that shrink towards the end (the automatic layout of large grids)
    ('jacobi2d', (70, 160), 2, dict(time_block=2, iterate=5, host_chunks=-5)),
    ('blur', (300, 60), 2, dict(time_block=1, iterate=3, host_chunks=3)),
    ('denoise2d', (64, 50), 2, dict(host_chunks=2)),
    ('heat3d', (40, 12, 40), 2, dict(time_block=2, iterate=4, host_chunks=3,
                                     options={'rows': 8})),
])
@pytest.mark.parametrize('edge_chunks', [None, 'natural'])
def test_slab_host_pipeline(tmp_path, name, extent, world, kwargs, edge_chunks):
  """soda_cuda_slab_run_host: every rank's own slices in host arrays, chunked
  H2D / passes / D2H per rank, the input ghosts from the neighbours' uploads
  (one exchange per call); bit-identical to the single-rank oracle, bytes
  outside the valid box untouched."""
  # edge_chunks: where the chunks that read ghost slices are computed - last
  # (the default of a callback transport) or in their natural place (what the
  # NCCL transport does on GPUs)
  run_case(tmp_path, name, extent, world, exchange_every=-1,
           edge_chunks=edge_chunks, **kwargs)


@pytest.mark.parametrize('name,extent,devices,kwargs', [
    ('jacobi2d', (70, 120), 3, dict(time_block=2, iterate=5)),
    ('blur', (300, 60), 2, dict(time_block=1, iterate=3)),
    ('denoise2d', (64, 50), 4, {}),
    ('heat3d', (40, 12, 40), 2, dict(time_block=2, iterate=4,
                                     options={'rows': 8})),
    ('jacobi2d', (40, 9), 4, dict(time_block=2, iterate=4)),  # thinner than halo
])
def test_one_process_several_devices(name, extent, devices, kwargs):
  """soda_cuda_multi_run_host (sodac --cuda-gpus): one process, the whole grid
  in host arrays, one slab per (emulated) device, ghosts read from the host
  arrays - and the same through the program-named entry point with
  opts.reserved[1]."""
  from soda_b200.codegen.cuda import launcher, multi_gpu
  from tests.emu import build_emu
  kwargs = dict(kwargs)
  time_block = kwargs.pop('time_block', None)
  options = kwargs.pop('options', None)
  st = common.stencil(name, **kwargs)
  prog = launcher.CudaProgram(
      build_emu.build_emu_library(st, time_block=time_block, options=options))
  inputs = common.make_inputs(st, extent, seed=9)
  want = common.oracle_outputs(st, inputs)
  for chunks in (1, 3):
    outputs = {n: np.full(extent[::-1], 77, dtype=d)
               for n, d in zip(prog.output_names, prog.output_dtypes)}
    multi_gpu.run_host_multi(prog, inputs, outputs, num_devices=devices,
                             opts=launcher.make_opts(host_chunks=chunks))
    common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)
  outputs = {n: np.full(extent[::-1], 77, dtype=d)
             for n, d in zip(prog.output_names, prog.output_dtypes)}
  prog.run_host(inputs, outputs, opts=launcher.make_opts(gpus=devices))
  common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)
