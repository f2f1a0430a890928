#!/usr/bin/env python3
"""ORACLE PINNING (test infrastructure; runs only where /root/reference exists).

Runs the REFERENCE's own golden-loop generator - ``print_test`` of
src/soda/codegen/frt/host.py:434-669, unmodified, imported from
/root/reference - to print its ``SODA_TEST_MAIN`` program for each shipped
stencil, compiles that program with g++ and lets it judge this repo's oracle:
the generated ``main`` initialises the inputs, calls ``soda::app::<app>`` (here
implemented by oracle/emit_cpp.py's restatement instead of the FPGA kernel),
recomputes every tensor with the reference's loops and compares element by
element (THRESHOLD=0, so floats must match exactly).

What comes from the reference: loop bounds, valid boxes, load/store index
arithmetic, initial values, the comparison and the PASS/FAIL verdict.
What cannot (the reference's ``haoda`` dependency is not installed): the text
of each C expression, which is printed by this repo's soda_b200.ir.CPrinter,
and the tensor DAG, which is built by this repo's front end and checked
against the reference separately (tests/golden/windows.json).

Output: oracle/_ref/<app>.cpp, oracle/_ref/<app>.exe and
tests/golden/reference_pin.log (committed) with one PASS/FAIL line per case.
"""
import contextlib
import io
import os
import subprocess
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, '_ref')

from tests.golden import make_reference_fixtures as fixtures  # noqa: E402


class MetaFmt:
  def __init__(self, fmt):
    self.fmt = fmt
  def __getitem__(self, key):
    return self.fmt % key


class CppPrinter:
  """Stand-in for haoda.util.CppPrinter (only what print_test uses)."""
  def __init__(self, out):
    self.out = out
    self.indent = 0
  def println(self, line='', indent=None):
    pad = self.indent if indent is None else indent
    self.out.write('  ' * pad + line + '\n')
  def printlns(self, *lines):
    for line in lines:
      if isinstance(line, str):
        self.println(line)
      else:
        for sub in line:
          self.println(sub)
  def do_scope(self):
    self.println('{')
    self.indent += 1
  def un_scope(self):
    self.indent -= 1
    self.println('}')
  @contextlib.contextmanager
  def if_(self, cond):
    self.println('if (%s)' % cond)
    self.do_scope()
    yield
    self.un_scope()
  @contextlib.contextmanager
  def for_(self, *parts):
    self.println('for (%s)' % '; '.join(parts))
    self.do_scope()
    yield
    self.un_scope()
  def print_func(self, name, params, suffix='', align=0):
    self.println('%s(%s)%s' % (name, ', '.join(params), suffix))


def load_reference_host():
  ref_core = fixtures.load_reference_core()
  from soda_b200 import ir
  haoda_ir = sys.modules['haoda.ir']
  haoda_ir.Ref = ir.Ref
  haoda_ir.make_var = ir.make_var
  haoda_util = sys.modules['haoda.util']
  haoda_util.MetaFmt = MetaFmt
  haoda_util.CppPrinter = CppPrinter
  sys.modules['haoda'].ir = haoda_ir
  sys.modules['haoda'].util = haoda_util
  import soda
  soda.core = ref_core
  sys.modules['soda'].core = ref_core
  import importlib
  host = importlib.import_module('soda.codegen.frt.host')
  return host, ref_core


def add_c_expr():
  """The reference calls node.c_expr (a haoda property); give our nodes one."""
  from soda_b200 import ir
  printer = ir.CPrinter(ref_printer=lambda ref: (_ for _ in ()).throw(
      AssertionError('unmutated load %s' % ref)),
                        min_name='ref_min', max_name='ref_max')
  ir.Node.c_expr = property(lambda self: printer(self))


PRELUDE = r'''
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstdbool>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <iomanip>
#include <iostream>
#include <memory>
#include <random>
#include <string>
#include <vector>
using std::clog; using std::endl; using std::string; using std::vector;
template <typename T> inline T ref_min(T a, T b) { return b < a ? b : a; }
template <typename T> inline T ref_max(T a, T b) { return a < b ? b : a; }
template <typename T> inline T soda_abs(T a) { return a < 0 ? T(-a) : a; }
extern "C" int soda_oracle_run(const void* const* inputs, void* const* outputs,
                               const int64_t* extent);
'''


def build_case(host, program, overrides, dims, sabotage=None):
  from soda_b200 import sodac
  from oracle import emit_cpp
  with open(os.path.join(ROOT, 'tests', 'src', program + '.soda')) as fp:
    stencil = sodac.compile_source(fp.read(), **overrides)
  for tensor in stencil.tensors.values():
    if tensor.st_ref is not None:
      tensor.st_ref.haoda_type = tensor.haoda_type
  app = stencil.app_name
  window_dims = [hi - lo + 1 for lo, hi in zip(
      *stencil.window_bounds[stencil.output_names[0]])]
  text = io.StringIO()
  text.write(PRELUDE)
  text.write('namespace soda { namespace app {\n')
  for d, size in enumerate(window_dims):
    text.write('constexpr int kStencilDim%d = %d;\n' % (d, size))
  params = []
  names = list(stencil.input_names) + list(stencil.output_names)
  types = list(stencil.input_types) + list(stencil.output_types)
  for name, t in zip(names, types):
    const = 'const ' if name in stencil.input_names else ''
    params += ['%s%s* var_%s_ptr' % (const, t.c_type, name),
               'const int32_t* var_%s_extent' % name,
               'const int32_t* var_%s_stride' % name,
               'const int32_t* var_%s_min' % name]
  params.append('const char* bitstream')
  text.write('// stands in for the FPGA kernel + tiling wrapper: this repo\'s '
             'oracle\n')
  text.write('int %s(%s) {\n' % (app, ', '.join(params)))
  text.write('  const void* ins[] = {%s};\n' % ', '.join(
      'var_%s_ptr' % n for n in stencil.input_names))
  text.write('  void* outs[] = {%s};\n' % ', '.join(
      'var_%s_ptr' % n for n in stencil.output_names))
  text.write('  int64_t extent[%d];\n' % stencil.dim)
  text.write('  for (int d = 0; d < %d; ++d) extent[d] = var_%s_extent[d];\n' %
             (stencil.dim, stencil.input_names[0]))
  text.write('  return soda_oracle_run(ins, outs, extent);\n}\n')
  text.write('} }  // namespace soda::app\n')
  host.print_test(CppPrinter(text), stencil)   # <- the reference's generator
  os.makedirs(os.path.join(OUT, 'include'), exist_ok=True)
  open(os.path.join(OUT, 'include', 'ap_int.h'), 'w').write(
      '// empty stand-in: the generated test main includes <ap_int.h>\n')
  tag = app + ''.join('_%s%s' % kv for kv in sorted(overrides.items()))
  if sabotage is not None:
    tag += '_canary'
  src = os.path.join(OUT, tag + '.cpp')
  with open(src, 'w') as fp:
    fp.write(text.getvalue())
  oracle_src = os.path.join(OUT, tag + '_oracle.cpp')
  with open(oracle_src, 'w') as fp:
    oracle_text = emit_cpp.emit(stencil)
    if sabotage is not None:  # canary: the judge must notice a wrong oracle
      assert sabotage[0] in oracle_text
      oracle_text = oracle_text.replace(*sabotage)
    fp.write(oracle_text)
  exe = os.path.join(OUT, tag + '.exe')
  subprocess.run(['g++', '-std=c++17', '-O1', '-ffp-contract=off',
                  '-DSODA_TEST_MAIN', '-I', os.path.join(OUT, 'include'), src,
                  oracle_src, '-o', exe], check=True)
  env = dict(os.environ, THRESHOLD='0')
  result = subprocess.run([exe, ''] + [str(d) for d in dims], env=env,
                          capture_output=True, text=True)
  verdict = 'PASS' if result.returncode == 0 and 'PASS' in result.stderr \
      else 'FAIL'
  return tag, verdict, result.stderr.strip().splitlines()[-1:]


def main():
  host, _ = load_reference_host()
  add_c_expr()
  cases = []
  programs = sorted(f[:-5] for f in os.listdir(os.path.join(ROOT, 'tests',
                                                           'src'))
                    if f.endswith('.soda'))
  for program in programs:
    cases.append((program, {}, []))        # the reference's default size
  cases += [
      ('blur', {'iterate': 2}, [2000, 40]),
      ('jacobi2d', {'iterate': 5}, [32, 50]),
      ('jacobi2d', {}, [32, 300]),
      ('seidel2d', {'iterate': 3}, [32, 41]),
      ('heat3d', {'iterate': 3}, [32, 32, 19]),
      ('jacobi3d', {}, [32, 32, 40]),
      ('denoise2d', {}, [32, 64]),
      ('denoise3d', {}, [32, 32, 17]),
      ('xcorr', {}, [480, 64]),
      ('erosion', {}, [480, 47]),
      ('sobel2d', {}, [32, 100]),
      ('contrast', {}, [480, 40]),
  ]
  lines = ['# produced by oracle/pin_against_reference.py: the reference\'s '
           'print_test (src/soda/codegen/frt/host.py:434-669) judging '
           'oracle/emit_cpp.py, THRESHOLD=0 (exact)']
  failed = 0
  for program, overrides, dims in cases:
    tag, verdict, tail = build_case(host, program, overrides, dims)
    failed += verdict != 'PASS'
    lines.append('%s %s dims=%s %s' % (verdict, tag, dims or 'default',
                                       ' '.join(tail)))
    print(lines[-1])
  # canaries: a one-ulp-ish / off-by-one change in the oracle must be caught
  for program, sabotage in (('jacobi2d', ('0.2f', '0.2000001f')),
                            ('blur', ('(q + (2))', '(q + (1))'))):
    tag, verdict, tail = build_case(host, program, {}, [], sabotage=sabotage)
    ok = verdict == 'FAIL'
    failed += not ok
    lines.append('%s %s (sabotaged oracle is rejected: %s)' %
                 ('PASS' if ok else 'FAIL', tag, verdict))
    print(lines[-1])
  with open(os.path.join(ROOT, 'tests', 'golden', 'reference_pin.log'),
            'w') as fp:
    fp.write('\n'.join(lines) + '\n')
  return 1 if failed else 0


if __name__ == '__main__':
  sys.exit(main())
