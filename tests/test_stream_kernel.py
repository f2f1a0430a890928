"""The reference's whole host data path, end to end, on the CPU:

  dense arrays --tiler--> bank buffers --<app>_kernel--> bank buffers
              --un-tiler--> dense arrays  ==  golden loops (inside valid boxes)

`<app>_kernel` is the stream function of the FPGA kernel with the reference's
port interface (oracle/stream_kernel.py; hls_kernel.py:62-66,88), the tiler and
un-tiler are oracle/stream_layout.py (frt/host.py:181-249, :340-427), the
golden loops oracle/emit_cpp.py (frt/host.py:556-624).  This is what the
reference's tests/test-cpp-host.sh checks with its generated code; here it pins
the stream-layout constants (stencil distance, output offset, bank partition,
burst alignment, halo replication between tiles) against the golden semantics
by an independent route.  The `gpu` test puts the CUDA pack / unpack kernels
on the two ends instead."""
import numpy as np
import pytest

from oracle import emit_cpp, stream_kernel
from oracle import stream_layout as oracle_layout
from soda_b200.codegen.cuda import stream_layout
from tests import common
from tests.test_stream_layout import to_oracle

CASES = [
    ('blur', (150, 12), dict(tile_size=[64])),
    ('blur', (200, 9), dict(tile_size=[64], dram_in='0.1', dram_out='0.1')),
    ('jacobi2d', (70, 11), {}),                       # iterate 2, tile 32
    ('jacobi2d', (40, 9), dict(tile_size=[16], dram_in='0.1.2.3',
                               dram_out='0.1.2.3', iterate=3)),
    ('sobel2d', (50, 10), dict(tile_size=[24])),
    # two inputs, one tile: with several tiles the reference's un-tiler takes
    # its loop bounds from the window of the FIRST input (f, a single point,
    # frt/host.py:352-356) and overwrites the two columns at every tile seam
    # with cells whose u-window left the tile - a quirk of the reference that
    # the codec reproduces (see test_multi_input_seam_quirk)
    ('denoise2d', (32, 12), {}),
    ('seidel2d', (64, 10), dict(iterate=1)),
    ('heat3d', (20, 19, 7), dict(tile_size=[8, 8], iterate=1)),
    ('jacobi3d', (40, 37, 6), dict(iterate=1)),       # tile 32 x 32
]


def run_data_path(st, extent, inputs, pack, unpack, kernel=None):
  layouts = {n: stream_layout.TensorLayout(st, n, extent)
             for n in st.input_names + st.output_names}
  first = layouts[st.input_names[0]]
  per_tile = extent[-1]
  for size in first.tile_size:
    per_tile *= size
  assert per_tile % first.elem_count_per_cycle == 0, 'pick an unpadded case'
  in_banks = {n: pack(layouts[n], inputs[n]) for n in st.input_names}
  out_banks = {
      n: [np.zeros(layouts[n].elems_per_bank,
                   dtype=inputs[st.input_names[0]].dtype if False else
                   np.dtype(common.golden.np_dtype(
                       st.output_stmts[i].haoda_type)))
          for _ in range(layouts[n].banks)]
      for i, n in enumerate(st.output_names)
  }
  cycles = stream_kernel.cycle_count(st, extent)
  for n in st.input_names + st.output_names:
    words = layouts[n].burst_width // layouts[n].elem_bits
    assert cycles * words <= layouts[n].elems_per_bank
  (kernel or stream_kernel.StreamKernel)(st).run(in_banks, out_banks, cycles)
  outputs = {}
  for i, n in enumerate(st.output_names):
    dense = np.full(extent[::-1], 77,
                    dtype=common.golden.np_dtype(st.output_stmts[i].haoda_type))
    outputs[n] = unpack(layouts[n], out_banks[n], dense)
  return outputs


@pytest.mark.parametrize('name,extent,overrides', CASES)
def test_tiler_kernel_untiler_equals_golden_loops(name, extent, overrides):
  st = common.stencil(name, **overrides)
  inputs = common.make_inputs(st, extent, seed=21)
  got = run_data_path(
      st, extent, inputs,
      pack=lambda layout, dense: oracle_layout.tile(to_oracle(layout, st), dense),
      unpack=lambda layout, banks, dense: oracle_layout.untile(
          to_oracle(layout, st), banks, dense))
  want = emit_cpp.Oracle(st).run(inputs)
  # the un-tiler writes the box of the window first input -> first output,
  # which contains every output's valid box; compare inside the valid boxes
  for out in st.output_names:
    index = common.box_index(st.valid_box(out, extent))
    a, b = got[out][index], want[out][index]
    assert a.size > 0
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), out


@pytest.mark.gpu
@pytest.mark.parametrize('name,extent,overrides', CASES[:5] + CASES[7:8])
def test_data_path_with_cuda_pack_and_unpack(name, extent, overrides):
  import torch
  st = common.stencil(name, **overrides)
  inputs = common.make_inputs(st, extent, seed=22)
  lib = stream_layout.LayoutLibrary()

  def signed(array):
    return array.view({2: np.int16, 4: np.int32}.get(array.itemsize,
                                                      array.dtype))

  def pack(layout, dense):
    banks = lib.pack(layout, torch.from_numpy(signed(dense).copy()).cuda())
    return [b.cpu().numpy().view(dense.dtype) for b in banks]

  def unpack(layout, banks, dense):
    device = torch.from_numpy(signed(dense).copy()).cuda()
    lib.unpack(layout, [torch.from_numpy(signed(b).copy()).cuda()
                        for b in banks], device)
    return device.cpu().numpy().view(dense.dtype)

  got = run_data_path(st, extent, inputs, pack, unpack)
  want = emit_cpp.Oracle(st).run(inputs)
  for out in st.output_names:
    index = common.box_index(st.valid_box(out, extent))
    assert np.array_equal(got[out][index].view(np.uint8),
                          want[out][index].view(np.uint8)), out


def test_multi_input_seam_quirk():
  """Documented behaviour of the reference, not of this repo: for denoise2d
  (inputs f, u) the un-tiler's bounds come from f's one-point window, so each
  later tile also writes its first columns, whose u-window crossed the tile
  boundary.  Everything else equals the golden loops."""
  st = common.stencil('denoise2d')
  extent = (60, 12)  # two tiles of 32 with stride 28
  inputs = common.make_inputs(st, extent, seed=21)
  got = run_data_path(
      st, extent, inputs,
      pack=lambda layout, dense: oracle_layout.tile(to_oracle(layout, st), dense),
      unpack=lambda layout, banks, dense: oracle_layout.untile(
          to_oracle(layout, st), banks, dense))
  want = emit_cpp.Oracle(st).run(inputs)
  index = common.box_index(st.valid_box('output', extent))
  differs = got['output'][index].view(np.uint32) != \
      want['output'][index].view(np.uint32)
  columns = sorted(set(np.argwhere(differs)[:, 1] + 2))  # box starts at x = 2
  assert columns == [28, 29]  # the seam: tile 1 starts at column 28
