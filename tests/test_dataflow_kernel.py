"""The reference-style dataflow kernel (modules, FIFOs, delay lines, burst
words; oracle/dataflow_kernel.py over the stand-in `hls_stream.h` / `ap_int.h`
of oracle/shim) on the reference's host data path:

  dense arrays --tiler--> bank buffers --<app>_kernel--> bank buffers
              --un-tiler--> dense arrays  ==  golden loops (inside valid boxes)

(the same harness as tests/test_stream_kernel.py, which runs the stream
function printed without any micro-architecture).  Structure checks pin
the module graph to the reference's documented example (README.md:127-156:
jacobi2d on a 2000-wide tile reads offsets {0, 1999, 2000, 2001, 4000} of its
input)."""
import numpy as np
import pytest

from oracle import dataflow_kernel, emit_cpp
from oracle import stream_layout as oracle_layout
from tests import common
from tests.test_stream_kernel import CASES, run_data_path
from tests.test_stream_layout import to_oracle

# unroll factors and banks beyond what the shipped programs declare
EXTRA = [
    ('jacobi2d', (70, 11), dict(unroll_factor=1)),
    ('jacobi2d', (70, 11), dict(unroll_factor=4, burst_width=256)),
    ('blur', (150, 12), dict(tile_size=[64], unroll_factor=8, burst_width=128)),
    ('jacobi2d', (40, 9), dict(tile_size=[16], dram_in='0.1', dram_out='0.1',
                               unroll_factor=4, burst_width=128, iterate=3)),
]


def _oracle_path(st, extent, inputs, kernel):
  return run_data_path(
      st, extent, inputs,
      pack=lambda layout, dense: oracle_layout.tile(to_oracle(layout, st), dense),
      unpack=lambda layout, banks, dense: oracle_layout.untile(
          to_oracle(layout, st), banks, dense),
      kernel=kernel)


def _cases():
  for name, extent, overrides in CASES + EXTRA:
    overrides = dict(overrides)
    banks = len(overrides.get('dram_in', '0').split('.'))
    if banks > 2 and 'unroll_factor' not in overrides:
      # a tensor's banks are dealt the elements of a cycle in turn: the unroll
      # factor has to be a multiple of the bank count (the shipped programs
      # declare unroll factor 2)
      overrides['unroll_factor'] = banks
    yield name, extent, overrides


@pytest.mark.parametrize('name,extent,overrides', list(_cases()))
def test_dataflow_kernel_equals_golden_loops(name, extent, overrides):
  st = common.stencil(name, **overrides)
  inputs = common.make_inputs(st, extent, seed=23)
  got = _oracle_path(st, extent, inputs, dataflow_kernel.DataflowKernel)
  want = emit_cpp.Oracle(st).run(inputs)
  for out in st.output_names:
    index = common.box_index(st.valid_box(out, extent))
    a, b = got[out][index], want[out][index]
    assert a.size > 0
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), out


def test_module_graph_of_the_documented_example():
  """README.md:127-156 of the reference: jacobi2d, tile 2000, unroll 1 keeps
  two lines of the input in its reuse buffers and reads five offsets."""
  st = common.stencil('jacobi2d', tile_size=[2000], unroll_factor=1, iterate=1)
  modules = dataflow_kernel.build_graph(st)
  forwards = [m for m in modules if m.kind == 'forward']
  assert sorted(int(m.name.rsplit('_', 1)[1]) for m in forwards) == \
      [0, 1999, 2000, 2001, 4000]
  assert sum(m.delay for m in forwards) == 4000
  assert [m.kind for m in modules].count('compute') == 1
  counts = dataflow_kernel.summary(st)
  assert counts['load'] == 1 and counts['store'] == 1
  # 5 taps + 4 chain links + load -> head + compute -> store
  assert counts['fifos'] == 5 + 4 + 1 + 1


def test_identical_modules_are_printed_once():
  st = common.stencil('jacobi2d', unroll_factor=4, burst_width=128)
  source = dataflow_kernel.emit(st)
  modules = dataflow_kernel.build_graph(st)
  definitions = source.count('\nvoid Module')
  assert definitions < len(modules)
  # the four processing elements of a stage share one definition
  assert source.count('// compute module') == 1


def test_shim_word_slicing(tmp_path):
  """`ap_uint<W>` of oracle/shim: range reads / writes at byte and non-byte
  positions, storage = W / 8 little-endian bytes."""
  import subprocess
  src = tmp_path / 't.cpp'
  src.write_text(r'''
#include <ap_int.h>
#include <hls_stream.h>
#include <cstdio>
int main() {
  ap_uint<512> w;
  static_assert(sizeof(w) == 64, "raw bytes");
  static_assert(sizeof(ap_uint<32>) == 4 && sizeof(ap_uint<16>) == 2, "");
  for (int i = 0; i < 16; ++i) w(i * 32 + 31, i * 32) = ap_uint<32>(0x1000u + i);
  const unsigned* raw = reinterpret_cast<const unsigned*>(&w);
  for (int i = 0; i < 16; ++i) if (raw[i] != 0x1000u + i) return 1;
  ap_uint<32> e = static_cast<ap_uint<32>>(w(95, 64));
  if (uint64_t(e) != 0x1002u) return 2;
  ap_uint<64> n;
  n(10, 5) = 0x2Bu;
  if (uint64_t(n) != (0x2Bull << 5)) return 3;
  if (uint64_t(n(10, 5)) != 0x2Bu) return 4;
  hls::stream<int> s("s");
  s.write(1); s.write(2);
  if (s.read() != 1 || s.empty() || s.read() != 2 || !s.empty()) return 5;
  std::puts("ok");
  return 0;
}''')
  exe = tmp_path / 't'
  subprocess.run(['g++', '-std=c++17', '-I', dataflow_kernel.SHIM_DIR,
                  str(src), '-o', str(exe)], check=True)
  assert subprocess.run([str(exe)], capture_output=True,
                        text=True).stdout.strip() == 'ok'
