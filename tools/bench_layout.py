#!/usr/bin/env python3
"""Throughput of the stream-data-layout pack / unpack kernels on cuda:0:
algorithmic bytes (one read + one write per stream element / per valid cell)
over the CUDA-event time, against the measured HBM copy peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from soda_b200 import sodac  # noqa: E402
from soda_b200.codegen.cuda import stream_layout  # noqa: E402


def main():
  import torch
  peak = 6535.1
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fp:
      peak = float(json.load(fp)['hbm_gbs'])
  except Exception:  # pylint: disable=broad-except
    pass
  lib = stream_layout.LayoutLibrary()
  cases = [
      ('jacobi2d', (16384, 16384), dict(tile_size=[2000]), torch.float32),
      ('jacobi2d', (16384, 16384),
       dict(tile_size=[2000], dram_in='0.1.2.3', dram_out='0.1.2.3'),
       torch.float32),
      ('blur', (16000, 16384), dict(tile_size=[2000]), torch.int16),
      ('heat3d', (512, 512, 512), dict(tile_size=[128, 128]), torch.float32),
  ]
  for name, extent, overrides, dtype in cases:
    with open(os.path.join(ROOT, 'tests', 'src', name + '.soda')) as fp:
      st = sodac.compile_source(fp.read(), **overrides)
    t_in = stream_layout.TensorLayout(st, st.input_names[0], extent)
    t_out = stream_layout.TensorLayout(st, st.output_names[0], extent)
    shape = tuple(extent[::-1])
    dense = torch.zeros(shape, dtype=dtype, device='cuda')
    banks = lib.pack(t_in, dense)
    out_banks = [torch.zeros(t_out.elems_per_bank, dtype=dtype, device='cuda')
                 for _ in range(t_out.banks)]

    def timed(fn, reps=5):
      for _ in range(2):
        fn()
      torch.cuda.synchronize()
      start = torch.cuda.Event(enable_timing=True)
      end = torch.cuda.Event(enable_timing=True)
      start.record()
      for _ in range(reps):
        fn()
      end.record()
      torch.cuda.synchronize()
      return start.elapsed_time(end) / reps

    stream = torch.cuda.current_stream().cuda_stream
    ms_pack = timed(lambda: lib.pack_device(
        t_in, dense.data_ptr(), [b.data_ptr() for b in banks], stream))
    ms_unpack = timed(lambda: lib.unpack_device(
        t_out, [b.data_ptr() for b in out_banks], dense.data_ptr(), stream))
    elem = dense.element_size()
    pack_bytes = 2 * t_in.elems_per_bank * t_in.banks * elem
    cells = 1
    for d in range(len(extent)):
      cells *= extent[d] - (t_out.window_dim[d] - 1)
    unpack_bytes = 2 * cells * elem
    print(json.dumps(dict(
        program=name, extent=extent, banks=t_in.banks, tiles=t_in.tile_count,
        pack_ms=ms_pack, pack_gbs=pack_bytes / ms_pack / 1e6,
        pack_frac=pack_bytes / ms_pack / 1e6 / peak,
        unpack_ms=ms_unpack, unpack_gbs=unpack_bytes / ms_unpack / 1e6,
        unpack_frac=unpack_bytes / ms_unpack / 1e6 / peak)), flush=True)
    del dense, banks, out_banks
    torch.cuda.empty_cache()


if __name__ == '__main__':
  main()
