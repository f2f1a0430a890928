"""1-D programs on the 2-D kernel template.

The reference's core is dimension-generic (reference: src/soda/core.py:858-926
compute windows for any ``dim``); the CUDA templates stream the last dimension
and vectorise dimension 0, so they need at least two.  A 1-D program
``b(0) = f(a(-1), a(0), a(1))`` over ``N`` cells is the 2-D program
``b(0, 0) = f(a(-1, 0), a(0, 0), a(1, 0))`` over an ``N x 1`` grid: the only
dimension becomes dimension 0 (lanes, shuffles), the streamed dimension has one
slice.  Values, valid range and iteration chain are unchanged; every strip of
the row is one (short-lived) warp, so this is the functional path for 1-D
programs, not a fast one.

The rewrite happens in the DSL itself, like the width lowering
(optimization/widths.py): the program is printed, every tensor reference gets
a second index 0, the first input gets a tile size, and the text is parsed
again.
"""
import re

from soda_b200 import util

LIFT_TILE = 32  # tile size printed for the new dimension 0 (no effect on results)


def lift_1d(stencil):
  """Returns the 2-D equivalent of a 1-D ``Stencil``."""
  from soda_b200 import sodac
  if stencil.dim != 1:
    raise util.InternalError('lift_1d expects a 1-D program')
  names = (list(stencil.input_names) + list(stencil.local_names) +
           list(stencil.output_names))
  text = str(stencil)
  pattern = re.compile(r'\b(%s)\(\s*([-+]?\d+)\s*\)' %
                       '|'.join(re.escape(n) for n in names))
  text = pattern.sub(lambda m: '%s(%s, 0)' % (m.group(1), m.group(2)), text)
  # the first input carries the tile size of the new dimension 0
  first = stencil.input_names[0]
  declaration = re.compile(r'^(input [^:\n]*:\s*%s)\s*$' % re.escape(first),
                           re.MULTILINE)
  text, count = declaration.subn(r'\1(%d, *)' % LIFT_TILE, text, count=1)
  if count != 1:
    raise util.InternalError('cannot find the declaration of input %s' % first)
  # computation reuse / inlining were applied before printing; the arithmetic
  # mode must survive the round trip
  lifted = sodac.compile_source(text,
                                math_precision=stencil.math_precision)
  if lifted.dim != 2:
    raise util.InternalError('lifting produced a %d-D program' % lifted.dim)
  return lifted
