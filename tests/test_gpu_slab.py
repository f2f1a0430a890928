"""The slab and multi-device entry points of the C ABI on a real GPU
(include/soda_cuda.h, "multi-GPU" section).  One visible device is enough for
the world-size-1 paths; the two-device cases run where the box has them.  Real
multi-rank runs (NCCL between processes) are checked by bench.py's parity
fields at N > 1 and by tools/multi_gpu_check.py."""
import numpy as np
import pytest
import torch

from soda_b200.codegen import cuda as cuda_backend
from soda_b200.codegen.cuda import launcher, multi_gpu
from tests import common

pytestmark = pytest.mark.gpu


def fresh_outputs(prog, extent):
  return {n: np.full(extent[::-1], 77, dtype=d)
          for n, d in zip(prog.output_names, prog.output_dtypes)}


@pytest.mark.parametrize('name,extent,kwargs', [
    ('jacobi2d', (1000, 900), dict(iterate=11, time_block=4)),
    ('blur', (2000, 512), dict(iterate=2)),
    ('denoise2d', (500, 300), {}),
    ('heat3d', (200, 40, 90), dict(iterate=5, time_block=2)),
])
def test_single_rank_slab_device_and_host_paths(name, extent, kwargs):
  kwargs = dict(kwargs)
  time_block = kwargs.pop('time_block', None)
  st = common.stencil(name, **kwargs)
  prog = cuda_backend.compile_stencil(st, time_block=time_block)
  inputs = common.make_inputs(st, extent, seed=21)
  want = common.oracle_outputs(st, inputs)
  device = torch.device('cuda', 0)
  runner = multi_gpu.SlabRunner(prog, extent, device, rank=0, world=1,
                                exchange_every=-1, host_chunks=3)
  assert runner.own == (0, extent[-1]) and len(runner.groups) == 1
  for tensor, iname in zip(runner.inputs, st.input_names):
    runner.view(tensor).copy_(torch.from_numpy(inputs[iname]))
  for tensor in runner.outputs:
    tensor.fill_(77)
  before = prog.launch_count()
  runner.run()
  torch.cuda.synchronize()
  assert prog.launch_count() - before == prog.num_passes
  got = {n: runner.view(t).cpu().numpy()
         for n, t in zip(st.output_names, runner.outputs)}
  common.assert_matches_oracle(st, extent, got, want, sentinel=77)
  outputs = fresh_outputs(prog, extent)
  runner.run_host(inputs, outputs)
  common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)
  runner.close()


@pytest.mark.parametrize('devices', [1, 2, 4])
def test_one_process_several_devices(devices):
  if torch.cuda.device_count() < devices:
    pytest.skip('needs %d visible devices' % devices)
  st = common.stencil('jacobi2d', iterate=11)
  prog = cuda_backend.compile_stencil(st, time_block=4)
  extent = (1000, 1800)
  inputs = common.make_inputs(st, extent, seed=22)
  want = common.oracle_outputs(st, inputs)
  outputs = fresh_outputs(prog, extent)
  multi_gpu.run_host_multi(prog, inputs, outputs, num_devices=devices,
                           opts=launcher.make_opts(host_chunks=3))
  common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)
  # the program-named entry point with opts.reserved[1] (sodac --cuda-gpus)
  outputs = fresh_outputs(prog, extent)
  prog.run_host(inputs, outputs, opts=launcher.make_opts(gpus=devices))
  common.assert_matches_oracle(st, extent, outputs, want, sentinel=77)


def test_pinned_host_buffers():
  st = common.stencil('jacobi2d')
  prog = cuda_backend.compile_stencil(st)
  extent = (700, 300)
  buf_in = launcher.HostBuffer(prog, extent[::-1], np.float32, 0)
  buf_out = launcher.HostBuffer(prog, extent[::-1], np.float32, 0)
  inputs = common.make_inputs(st, extent, seed=23)
  buf_in.array[...] = inputs['t1']
  buf_out.array[...] = 77
  prog.run_host({'t1': buf_in.array}, {'t0': buf_out.array})
  common.assert_matches_oracle(st, extent, {'t0': buf_out.array},
                               common.oracle_outputs(st, inputs), sentinel=77)
  buf_in.close()
  buf_out.close()


def test_too_many_devices_is_an_error():
  st = common.stencil('jacobi2d')
  prog = cuda_backend.compile_stencil(st)
  inputs = common.make_inputs(st, (100, 50), seed=1)
  with pytest.raises(launcher.SodaCudaError) as err:
    multi_gpu.run_host_multi(prog, inputs, num_devices=64)
  assert 'devices' in str(err.value)
