#!/usr/bin/env python3
"""Generates tests/golden/windows.json by calling the REFERENCE's own analysis
functions (src/soda/core.py: get_overall_stencil_window, get_stencil_distance,
get_stencil_window_offset, get_stencil_dim; src/soda/util.py) on the tensor
DAGs of the shipped programs.

Runs only in the build container, where /root/reference exists.  The
reference's third-party imports (haoda, pulp, toposort, cached_property, tapa)
are not installed, so they are replaced by empty stand-in modules: the four
functions above are pure Python over duck-typed tensors (name, st_idx, parents,
ld_indices) and never touch them.  Tensors are built by this repo's front end;
what the fixture pins is that our window analyses give the reference's numbers
for the same DAG.
"""
import importlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = '/root/reference/src'


def stub(name, **attrs):
  module = types.ModuleType(name)
  module.__dict__.update(attrs)
  sys.modules[name] = module
  return module


def load_reference_core():
  class Anything:
    def __getattr__(self, name):
      return Anything()
    def __call__(self, *a, **k):
      return Anything()
    def __mro_entries__(self, bases):
      return (object,)
  any_ = Anything()
  stub('cached_property', cached_property=lambda f: property(f))
  stub('pulp', PULP_CBC_CMD=lambda **k: None)
  stub('toposort')
  haoda = stub('haoda', ir=any_, util=any_)
  stub('haoda.ir', arithmetic=any_, GRAMMAR='', Node=object)
  stub('haoda.ir.arithmetic', base=any_)
  stub('haoda.util')
  stub('haoda.backend', xilinx=any_)
  stub('tapa')
  sys.path.insert(0, REFERENCE)
  # only soda.util and soda.core's free functions are needed; keep the other
  # reference modules from importing
  for name in ('soda.tensor', 'soda.dataflow', 'soda.visitor',
               'soda.optimization', 'soda.optimization.cluster',
               'soda.optimization.computation_reuse',
               'soda.optimization.inline', 'soda.grammar', 'soda.mutator'):
    stub(name, cluster=any_, computation_reuse=any_, inline=any_,
         InputStmt=object, LocalStmt=object, OutputStmt=object,
         ParamStmt=object, Tensor=object)
  import soda  # the reference package
  soda.tensor = sys.modules['soda.tensor']
  soda.grammar = sys.modules['soda.grammar']
  soda.util = importlib.import_module('soda.util')
  return importlib.import_module('soda.core')


def main():
  ref_core = load_reference_core()
  from soda_b200 import sodac  # noqa: E402
  cases = []
  src_dir = os.path.join(ROOT, 'tests', 'src')
  programs = sorted(f[:-5] for f in os.listdir(src_dir) if f.endswith('.soda'))
  variants = [(p, {}) for p in programs] + [('blur', {'iterate': 2}),
                                            ('jacobi2d', {'iterate': 5}),
                                            ('heat3d', {'iterate': 3})]
  for program, overrides in variants:
    with open(os.path.join(src_dir, program + '.soda')) as fp:
      stencil = sodac.compile_source(fp.read(), **overrides)
    inputs = [stencil.tensors[n] for n in stencil.input_names]
    tensors = {}
    for tensor in stencil.chronological_tensors:
      if tensor.is_input():
        continue
      ref_core._overall_stencil_window_cache.clear()
      window = ref_core.get_overall_stencil_window(inputs, tensor)
      tensors[tensor.name] = {
          'window': [list(p) for p in window],
          'distance': ref_core.get_stencil_distance(window, stencil.tile_size),
          'offset': list(ref_core.get_stencil_window_offset(window)),
          'dim': list(ref_core.get_stencil_dim(window)),
      }
    out = stencil.tensors[stencil.output_names[0]]
    ref_core._overall_stencil_window_cache.clear()
    window = ref_core.get_overall_stencil_window(inputs, out)
    distance = ref_core.get_stencil_distance(window, stencil.tile_size)
    offset = distance - sys.modules['soda.util'].serialize(
        ref_core.get_stencil_window_offset(window), stencil.tile_size)
    cases.append({
        'program': program,
        'overrides': overrides,
        'tensors': tensors,
        'stencil_distance': max(distance, offset),
    })
  with open(os.path.join(HERE, 'windows.json'), 'w') as fp:
    json.dump({'generated_by': 'tests/golden/make_reference_fixtures.py',
               'reference_functions': 'src/soda/core.py:858-926',
               'cases': cases}, fp, separators=(',', ':'))
  print('wrote', len(cases), 'cases')


if __name__ == '__main__':
  main()
