"""The reference's stream data layout (tiler / un-tiler of the generated host
wrapper, src/soda/codegen/frt/host.py:112-249, :340-427).

CPU tests pin the oracle restatement (oracle/stream_layout.py) against the
layouts drawn in the reference's docs/data-layout.md and check the host-side
constants and the C ABI; the `gpu` tests compare the CUDA pack / unpack kernels
with the oracle bit for bit."""
import ctypes

import numpy as np
import pytest

from oracle import stream_layout as oracle_layout
from soda_b200.codegen.cuda import build as cuda_build
from soda_b200.codegen.cuda import stream_layout
from tests import common

BLUR_3X3 = '''
kernel: blur3
burst width: 16
unroll factor: 1
iterate: 1
input dram {banks} uint16: input(100, *)
output dram {banks} uint16: out(0, 0) = (input(0, 0) + input(1, 0) + input(2, 0) + input(0, 1) + input(1, 1) + input(2, 1) + input(0, 2) + input(1, 2) + input(2, 2)) / 9
'''


def blur3(banks=1):
  from soda_b200 import sodac
  spec = '.'.join(str(b) for b in range(banks))
  return sodac.compile_source(BLUR_3X3.format(banks=spec))


def to_oracle(t: stream_layout.TensorLayout, st) -> oracle_layout.StreamLayout:
  """The oracle's description of the same tensor, from the host module's
  fields (the formulas that derive buffer sizes live in both and are compared
  below)."""
  first_in = st.input_stmts[0]
  ref = st.output_stmts[0] if t.is_output else first_in
  return oracle_layout.StreamLayout(
      extent=t.extent, tile_size=t.tile_size, stencil_dim=t.stencil_dim,
      stencil_distance=t.stencil_distance, window_offset=t.window_offset,
      window_dim=t.window_dim, stencil_offset=t.stencil_offset,
      elem_bits=t.elem_bits, banks=t.banks, burst_width=t.burst_width,
      ref_elem_bits=ref.haoda_type.width_in_bits, ref_banks=len(ref.dram),
      first_input_elem_bits=first_in.haoda_type.width_in_bits,
      first_input_banks=len(first_in.dram), produce_offset=t.produce_offset)


def image(width, height):
  """Pixel (p, q) holds 1000 q + p + 1: every cell is recognisable and none
  equals the void marker 0."""
  q, p = np.mgrid[0:height, 0:width]
  return (1000 * q + p + 1).astype(np.uint16)


# ---- golden vectors: docs/data-layout.md ----------------------------------------

def test_doc_single_bank_single_tile():
  """docs/data-layout.md:18-50: 3x3 blur on 100x100, stencil distance 202; the
  input stream is the image followed by 202 void elements, the output stream
  starts with 202 void elements and pixel (0, 0) follows."""
  st = blur3()
  t_in = stream_layout.TensorLayout(st, 'input', (100, 100))
  t_out = stream_layout.TensorLayout(st, 'out', (100, 100))
  assert t_in.stencil_distance == 202 and t_out.stencil_offset == 202
  assert t_in.elems_per_bank == 100 * 100 + 202
  img = image(100, 100)
  (stream,) = oracle_layout.tile(to_oracle(t_in, st), img)
  assert np.array_equal(stream[:10000], img.reshape(-1))
  assert np.all(stream[10000:] == 0) and stream.size == 10202
  # output: what sits 202 elements into the stream is output pixel (0, 0)
  out_stream = np.arange(10202, dtype=np.uint16)
  out = np.full((100, 100), 7, dtype=np.uint16)
  oracle_layout.untile(to_oracle(t_out, st), [out_stream], out)
  assert out[0, 0] == 202 and out[0, 97] == 202 + 97
  assert out[1, 0] == 202 + 100 and out[97, 97] == 202 + 97 * 100 + 97
  # the halo of the output is not written (docs/data-layout.md:15-16)
  assert np.all(out[:, 98:] == 7) and np.all(out[98:, :] == 7)


def test_doc_two_banks_single_tile():
  """docs/data-layout.md:62-112: cyclic partition over the banks."""
  st = blur3(banks=2)
  t_in = stream_layout.TensorLayout(st, 'input', (100, 100))
  img = image(100, 100)
  bank0, bank1 = oracle_layout.tile(to_oracle(t_in, st), img)
  assert bank0.size == bank1.size == 5101
  # bank 0: (0,0) (2,0) ... (98,0) (0,1) ...; bank 1: (1,0) (3,0) ...
  assert list(bank0[:3]) == [img[0, 0], img[0, 2], img[0, 4]]
  assert bank0[49] == img[0, 98] and bank0[50] == img[1, 0]
  assert list(bank1[:3]) == [img[0, 1], img[0, 3], img[0, 5]]
  assert bank1[49] == img[0, 99] and bank1[4999] == img[99, 99]
  assert np.all(bank0[5000:] == 0) and np.all(bank1[5000:] == 0)
  t_out = stream_layout.TensorLayout(st, 'out', (100, 100))
  out = np.zeros((100, 100), dtype=np.uint16)
  streams = [np.arange(0, 10202, 2, dtype=np.uint16),
             np.arange(1, 10202, 2, dtype=np.uint16)]
  oracle_layout.untile(to_oracle(t_out, st), streams, out)
  # output bank 0 row 2: void, (0,0), (2,0) ...; bank 1: void, (1,0), (3,0) ...
  assert out[0, 0] == 202 and out[0, 1] == 203 and out[0, 2] == 204


def test_doc_single_bank_multi_tile():
  """docs/data-layout.md:128-178: a 150x150 image on a (100, *) kernel: the
  second tile starts at column 98 and is padded to the tile width."""
  st = blur3()
  t_in = stream_layout.TensorLayout(st, 'input', (150, 150))
  assert t_in.tile_count == [2]
  img = image(150, 150)
  (stream,) = oracle_layout.tile(to_oracle(t_in, st), img)
  assert stream.size == 2 * 150 * 100 + 202
  tile0 = stream[:15000].reshape(150, 100)
  tile1 = stream[15000:30000].reshape(150, 100)
  assert np.array_equal(tile0, img[:, :100])
  assert np.array_equal(tile1[:, :52], img[:, 98:150])
  assert np.all(tile1[:, 52:] == 0) and np.all(stream[30000:] == 0)
  # output: tile 0 holds columns 0..97, tile 1 columns 98..147
  t_out = stream_layout.TensorLayout(st, 'out', (150, 150))
  out = np.zeros((150, 150), dtype=np.uint32)
  oracle_layout.untile(to_oracle(t_out, st), [np.arange(30202, dtype=np.uint32)],
                       out)
  assert out[0, 0] == 202 and out[0, 97] == 202 + 97
  assert out[0, 98] == 15000 + 202 and out[0, 147] == 15000 + 202 + 49
  assert out[147, 147] == 15000 + 202 + 147 * 100 + 49
  assert np.all(out[:, 148:] == 0) and np.all(out[148:, :] == 0)


def test_doc_two_banks_multi_tile():
  """docs/data-layout.md:186-262."""
  st = blur3(banks=2)
  t_in = stream_layout.TensorLayout(st, 'input', (150, 150))
  img = image(150, 150)
  bank0, bank1 = oracle_layout.tile(to_oracle(t_in, st), img)
  # second tile, first row: (98,0) (100,0) ... in bank 0, (99,0) (101,0) ...
  assert list(bank0[7500:7503]) == [img[0, 98], img[0, 100], img[0, 102]]
  assert list(bank1[7500:7503]) == [img[0, 99], img[0, 101], img[0, 103]]
  assert bank0[7500 + 25] == img[0, 148] and bank0[7500 + 26] == 0


# ---- host-side constants and the C ABI ----------------------------------------------

CONFIGS = [
    ('blur', (2100, 9), dict(tile_size=[2000]), {}),
    ('blur', (300, 17), dict(tile_size=[64], dram_in='0.1', dram_out='2.3'),
     dict(burst_width=64)),
    ('jacobi2d', (100, 12), {}, {}),
    ('jacobi2d', (70, 9), dict(tile_size=[32], dram_in='0.1.2.3',
                               dram_out='0.1.2.3'), {}),
    ('sobel2d', (50, 11), dict(tile_size=[20]), {}),
    ('heat3d', (40, 37, 6), {}, {}),
    ('jacobi3d', (21, 19, 5), dict(tile_size=[8, 9], dram_out='0.1'), {}),
    ('denoise3d', (40, 33, 7), {}, {}),
]


@pytest.mark.parametrize('name,extent,overrides,kwargs', CONFIGS)
def test_layout_constants_match_the_oracle(name, extent, overrides, kwargs):
  st = common.stencil(name, **overrides)
  lib = stream_layout.LayoutLibrary()
  for tensor in st.input_names + st.output_names:
    t = stream_layout.TensorLayout(st, tensor, extent, **kwargs)
    o = to_oracle(t, st)
    assert t.tile_count == o.tile_count
    assert t.elem_count_per_cycle == o.elem_count_per_cycle
    assert t.elem_count_aligned_per_tile == o.elem_count_aligned_per_tile
    assert t.elems_per_bank == o.elems_per_bank
    # arithmetic only: no kernel is launched, works without a GPU
    assert lib.bank_elems(t) == o.elems_per_bank


def test_library_exports_every_declared_symbol():
  lib = ctypes.CDLL(cuda_build.build_layout_library())
  for symbol in stream_layout.LayoutLibrary.SYMBOLS:
    assert hasattr(lib, symbol), symbol
  assert ctypes.sizeof(stream_layout.CLayout) == 144  # sizeof(soda_stream_layout)


def test_bad_arguments_are_reported():
  st = common.stencil('jacobi2d')
  lib = stream_layout.LayoutLibrary()
  t = stream_layout.TensorLayout(st, 't1', (64, 8))
  t.banks = 0
  with pytest.raises(stream_layout.SodaLayoutError):
    lib.bank_elems(t)
  with pytest.raises(Exception):
    stream_layout.TensorLayout(st, 'nope', (64, 8))
  with pytest.raises(Exception):
    stream_layout.TensorLayout(st, 't1', (2, 8))  # narrower than the window
  # outputs slower than the first input: the reference's tiles would overlap
  lopsided = common.stencil('jacobi2d', dram_in='0.1.2.3')
  with pytest.raises(Exception):
    stream_layout.TensorLayout(lopsided, 't0', (64, 8))


# ---- the CUDA kernels against the oracle -----------------------------------------------

def _np_dtype(bits):
  return {8: np.uint8, 16: np.uint16, 32: np.uint32, 64: np.uint64}[bits]


@pytest.mark.gpu
@pytest.mark.parametrize('name,extent,overrides,kwargs', CONFIGS)
def test_pack_and_unpack_match_the_oracle(name, extent, overrides, kwargs):
  import torch
  st = common.stencil(name, **overrides)
  lib = stream_layout.LayoutLibrary()
  rng = np.random.default_rng(7)
  shape = tuple(extent[::-1])
  before = lib.launch_count()
  for tensor in st.input_names:
    t = stream_layout.TensorLayout(st, tensor, extent, **kwargs)
    dtype = _np_dtype(t.elem_bits)
    dense = rng.integers(1, 60000, shape).astype(dtype)
    want = oracle_layout.tile(to_oracle(t, st), dense)
    view = dense.view(np.int16 if t.elem_bits == 16 else
                      np.int32 if t.elem_bits == 32 else dense.dtype)
    got = lib.pack(t, torch.from_numpy(view.copy()).cuda())
    torch.cuda.synchronize()
    assert len(got) == t.banks
    for g, w in zip(got, want):
      assert np.array_equal(g.cpu().numpy().view(dtype), w), tensor
  for tensor in st.output_names:
    t = stream_layout.TensorLayout(st, tensor, extent, **kwargs)
    dtype = _np_dtype(t.elem_bits)
    signed = np.int16 if t.elem_bits == 16 else np.int32
    banks = [rng.integers(1, 60000, t.elems_per_bank).astype(dtype)
             for _ in range(t.banks)]
    want = oracle_layout.untile(to_oracle(t, st), banks,
                                np.full(shape, 77, dtype=dtype))
    dense = torch.full(shape, 77, dtype=torch.int16 if t.elem_bits == 16
                       else torch.int32).cuda()
    lib.unpack(t, [torch.from_numpy(b.view(signed).copy()).cuda()
                   for b in banks], dense)
    torch.cuda.synchronize()
    assert np.array_equal(dense.cpu().numpy().view(dtype), want), tensor
  assert lib.launch_count() - before == len(st.input_names) + len(
      st.output_names)


@pytest.mark.gpu
def test_large_grid_round_trip_property():
  """Size-independent property at a BASELINE-sized grid (16384 x 4096 fp32,
  tiles of 2000, 4 banks): packing the array of linear indices yields a stream
  in which every non-void element names its own source cell, and un-tiling a
  stream of linear stream offsets yields, in every valid cell, the offset the
  reference's formula gives."""
  import torch
  st = common.stencil('jacobi2d', tile_size=[2000], dram_in='0.1.2.3',
                      dram_out='0.1.2.3')
  extent = (16384, 4096)
  lib = stream_layout.LayoutLibrary()
  t_in = stream_layout.TensorLayout(st, 't1', extent)
  dense = torch.arange(1, extent[0] * extent[1] + 1, dtype=torch.int32,
                       device='cuda').reshape(extent[1], extent[0])
  banks = [b.cpu().numpy() for b in lib.pack(t_in, dense)]
  stream = np.stack(banks, axis=1).reshape(-1)  # cyclic partition undone
  aligned = t_in.elem_count_aligned_per_tile
  stride = t_in.tile_size[0] - t_in.stencil_dim[0] + 1
  rng = np.random.default_rng(3)
  for offset in rng.integers(0, stream.size, 2000):
    tile, off = divmod(int(offset), aligned)
    i, j = off % t_in.tile_size[0], off // t_in.tile_size[0]
    void = tile >= t_in.tile_count[0] or j >= extent[1] or \
        i >= (extent[0] - stride * tile if tile == t_in.tile_count[0] - 1
              else t_in.tile_size[0])
    want = 0 if void else j * extent[0] + tile * stride + i + 1
    assert stream[offset] == want, offset
  # every cell of the grid is in the stream at least once
  assert np.unique(stream).size == extent[0] * extent[1] + 1

  t_out = stream_layout.TensorLayout(st, 't0', extent)
  out_banks = [torch.arange(b, t_out.elems_per_bank * 4, 4, dtype=torch.int32,
                            device='cuda') for b in range(4)]
  out = torch.full((extent[1], extent[0]), -1, dtype=torch.int32, device='cuda')
  lib.unpack(t_out, out_banks, out)
  got = out.cpu().numpy()
  lo = t_out.window_offset
  cut = [t_out.window_dim[d] - 1 - lo[d] for d in range(2)]
  for _ in range(2000):
    p = int(rng.integers(0, extent[0]))
    q = int(rng.integers(0, extent[1]))
    valid = lo[0] <= p < extent[0] - cut[0] and lo[1] <= q < extent[1] - cut[1]
    if not valid:
      assert got[q, p] == -1
      continue
    tile = min((p - lo[0]) // stride, t_out.tile_count[0] - 1)
    want = tile * t_out.elem_count_aligned_per_tile + \
        q * t_out.tile_size[0] + (p - tile * stride) + t_out.stencil_offset
    assert got[q, p] == want, (p, q)
