"""Seeded random SODA programs for differential testing (test infrastructure).

``program(seed)`` returns (source text, Stencil overrides, extent, backend
kwargs).  The programs cover what the planner has to get right and the
hand-picked tests/src programs only sample: arbitrary tap offsets in every
dimension (asymmetric windows), off-centre store indices, DAGs of one to three
local stages with fan-out, several inputs, every element-type family (fp32,
binary16, 8/16/32-bit integers, narrow widths), iterate / time-block
combinations, packed and scalar arithmetic.

Values stay in range by construction: float stages are convex-ish combinations
(coefficients of magnitude <= 0.5), integer stages add a few taps and divide
or mask, so neither overflow nor NaN can make a bit-exact comparison vacuous
or ill-defined (signed overflow is undefined in the C++ evaluators).
"""
import numpy as np

FLOAT_FAMILIES = ('float', 'half')
INT_FAMILIES = ('int16', 'int32', 'uint16', 'uint8', 'uint6')


def _idx(rng, dim, reach, store=None):
  idx = tuple(int(rng.integers(-r, r + 1)) for r in reach[:dim])
  if store is not None:
    idx = tuple(a + b for a, b in zip(idx, store))
  return idx


def _ref(name, idx):
  return '%s(%s)' % (name, ', '.join(str(i) for i in idx))


def _float_expr(rng, family, sources, dim, reach, taps, store):
  """Sum of weighted taps; literals typed so that a half program stays a half
  program (``half(0.25f)``) or deliberately mixes in float arithmetic."""
  terms = []
  pure = rng.random() < 0.6  # only element-type operands: packable
  for k in range(taps):
    name = sources[int(rng.integers(len(sources)))]
    # the first tap sits on the stored cell: every stage's window contains
    # offset 0 in every dimension, otherwise a consumer's valid box would
    # index the stage outside its array (the reference's golden loops would
    # read out of bounds; such programs have no defined result)
    ref = _ref(name, store if k == 0 else _idx(rng, dim, reach, store))
    weight = float(rng.integers(1, 9)) / 16.0
    if family == 'half':
      lit = 'half(%sf)' % weight if pure or rng.random() < 0.5 else '%sf' % weight
    else:
      lit = '%sf' % weight
    kind = rng.integers(4)
    if kind == 0:
      terms.append(ref)
    elif kind == 1:
      terms.append('%s * %s' % (ref, lit))
    elif kind == 2:
      other = _ref(sources[int(rng.integers(len(sources)))],
                   _idx(rng, dim, reach, store))
      terms.append('%s * %s * %s' % (ref, other, lit))
    else:
      terms.append('(%s - %s)' % (lit, ref))
  expr = terms[0]
  for term in terms[1:]:
    expr += (' + ' if rng.random() < 0.7 else ' - ') + term
  scale = '%sf' % (1.0 / (taps + 1)) if family == 'float' else \
      'half(%sf)' % (1.0 / (taps + 1))
  expr = '(%s) * %s' % (expr, scale)
  if not pure and rng.random() < 0.5:
    a = _ref(sources[0], _idx(rng, dim, reach, store))
    expr = 'max(min(%s, %s), %s - 1) / (2 + %s * %s)' % (expr, a, a, a, a)
  return expr


def _int_expr(rng, sources, dim, reach, taps, store):
  terms = [_ref(sources[int(rng.integers(len(sources)))],
                store if k == 0 else _idx(rng, dim, reach, store))
           for k in range(taps)]
  expr = terms[0]
  for term in terms[1:]:
    expr += (' + ' if rng.random() < 0.6 else ' - ') + term
  # every stage ends with a range-reducing operation, so values cannot grow
  # over stages and iterations towards 2^31
  kind = rng.integers(5)
  if kind == 0:
    return '(%s) / %d' % (expr, int(rng.integers(5, 9)))
  if kind == 1:
    return '((%s) * %d) %% %d' % (expr, int(rng.integers(2, 5)),
                                 int(rng.integers(3, 200)))
  if kind == 2:
    return 'abs(%s) & %d' % (expr, int(rng.integers(15, 1024)))
  if kind == 3:
    a = terms[0]
    return '(max(%s, %s) - min(%s, 7) + (%s > %s) * 3) %% 3001' % (
        expr, a, expr, a, terms[-1])
  return '((%s) ^ (%s & 5) | 1) %% 2003' % (expr, terms[0])


def program(seed: int, hard: bool = False):
  """``hard``: wider windows (up to 4 cells in dimension 0), up to four local
  stages whose types differ from the tensors' (a double stage in a half
  program, an int32 stage between uint8 ones), ``let`` bindings, three inputs,
  double and 64-bit integer tensors.  The test-suite pins the plain seeds;
  tools/random_sweep.py explores the hard ones."""
  rng = np.random.default_rng((5000 if hard else 1000) + seed)
  dim = 2 if rng.random() < 0.65 else 3
  floating = rng.random() < 0.55
  float_families = FLOAT_FAMILIES + (('double',) if hard else ())
  int_families = INT_FAMILIES + (('int64',) if hard else ())
  family = (float_families if floating else int_families)[int(
      rng.integers(len(float_families if floating else int_families)))]
  # reach per dimension: dimension 0 costs shuffles, dimension 1 shared memory
  # (3-D) or register rows (2-D), the last one window depth
  reach = [int(rng.integers(0, 3)) for _ in range(dim)]
  if hard:
    reach[0] = int(rng.integers(0, 5))
    if dim == 2:
      reach[1] = int(rng.integers(0, 4))
  if not any(reach):
    reach[int(rng.integers(dim))] = 1
  num_inputs = 1 if rng.random() < 0.7 else 2
  if hard and rng.random() < 0.2:
    num_inputs = 3
  num_locals = int(rng.integers(0, 5 if hard else 3))
  iterate = 1
  if num_inputs == 1 and rng.random() < 0.6:
    iterate = int(rng.integers(2, 4))
  tile = ', '.join(['32'] * (dim - 1) + ['*'])

  lines = ['kernel: rnd%d' % seed, 'burst width: 64', 'unroll factor: 2',
           'iterate: %d' % iterate]
  inputs = ['in%d' % i for i in range(num_inputs)]
  for i, name in enumerate(inputs):
    lines.append('input dram %d %s: %s%s' %
                 (i, family, name, '(%s)' % tile if i == 0 else ''))
  sources = list(inputs)
  for k in range(num_locals):
    name = 'loc%d' % k
    taps = int(rng.integers(2, 5))
    store = _idx(rng, dim, [1] * dim) if rng.random() < 0.25 else (0,) * dim
    local_family = family
    if hard and rng.random() < 0.4:
      pool = float_families if floating else ('int16', 'int32', 'uint16',
                                              'int64')
      local_family = pool[int(rng.integers(len(pool)))]
    if floating:
      expr = _float_expr(rng, local_family, sources, dim, reach, taps, store)
    else:
      expr = _int_expr(rng, sources, dim, reach, taps, store)
    if hard and rng.random() < 0.4:
      # a typed and an untyped let, used twice
      first = _ref(sources[int(rng.integers(len(sources)))], store)
      other = _ref(sources[0], _idx(rng, dim, reach, store))
      if floating:
        lets = ('\n  %s t = %s * 0.5f\n  u = t - %s * t\n  ' %
                (local_family, first, other))
        expr = '(%s) * 0.5f + u * 0.25f - t * u * 0.125f' % expr
      else:
        lets = ('\n  int32 t = (%s + %s) / 2\n  u = t - %s\n  ' %
                (first, other, first))
        expr = '((%s) + (t & 255) - (u %% 7)) %% 4093' % expr
      lines.append('local %s:%s%s = %s' % (local_family, lets,
                                          _ref(name, store), expr))
    else:
      lines.append('local %s: %s = %s' % (local_family, _ref(name, store),
                                          expr))
    sources.append(name)
  taps = int(rng.integers(2, 6))
  # the output reads the newest stages preferentially so that the DAG is deep
  out_sources = sources[-2:] + [inputs[0]]
  origin = (0,) * dim
  if floating:
    expr = _float_expr(rng, family, out_sources, dim, reach, taps, origin)
  else:
    expr = _int_expr(rng, out_sources, dim, reach, taps, origin)
  lines.append('output dram %d %s: %s = %s' %
               (num_inputs, family, _ref('out', (0,) * dim), expr))
  text = '\n'.join(lines) + '\n'

  if dim == 2:
    extent = (int(rng.integers(60, 330)), int(rng.integers(30, 50)))
  else:
    extent = (int(rng.integers(50, 150)), int(rng.integers(30, 44)),
              int(rng.integers(26, 34)))
  kwargs = {}
  if iterate > 1:
    kwargs['time_block'] = int(rng.integers(1, iterate + 1))
  options = {}
  if rng.random() < 0.25:
    options['no_pack'] = True
  if dim == 2 and rng.random() < 0.2:
    options['no_pipeline'] = True
  if dim == 3 and rng.random() < 0.5:
    options['rows'] = 8
    options['cy'] = int(rng.choice([1, 2]))
  if options:
    kwargs['options'] = options
  return text, extent, kwargs


def inputs_for(st, extent, seed):
  """Float inputs in [0, 1), integers small enough that a handful of taps and
  a multiplication by < 5 stay far from 2^31 (and 8-bit cells keep their whole
  range)."""
  from oracle import golden
  rng = np.random.default_rng(seed)
  shape = tuple(extent[::-1])
  result = {}
  for stmt in st.input_stmts:
    dtype = golden.np_dtype(stmt.haoda_type)
    if stmt.haoda_type.is_float:
      result[stmt.name] = rng.random(shape, dtype=np.float32).astype(dtype)
    elif np.dtype(dtype).itemsize == 1:
      result[stmt.name] = rng.integers(0, 256, shape).astype(dtype)
    elif stmt.haoda_type.is_signed:
      result[stmt.name] = rng.integers(-1024, 1024, shape).astype(dtype)
    else:
      result[stmt.name] = rng.integers(0, 2048, shape).astype(dtype)
  return result


def sum_program(seed: int):
  """A one-stage sum of 4-11 taps drawn from a small box: the shape
  computation reuse rewrites.  Returns (text, extent, number of taps).  Integer
  programs make the rewrite checkable exactly (modular addition is associative
  and commutative), float ones within rounding."""
  rng = np.random.default_rng(7000 + seed)
  dim = 2 if rng.random() < 0.7 else 3
  t = ('int32', 'int16', 'float', 'uint16')[int(rng.integers(4))]
  box = [int(rng.integers(1, 3)) for _ in range(dim)]
  points = {(0,) * dim}
  capacity = 1
  for b in box:
    capacity *= 2 * b + 1
  taps = min(int(rng.integers(4, 12)), capacity)
  while len(points) < taps:
    points.add(tuple(int(rng.integers(-b, b + 1)) for b in box))
  refs = ' + '.join(_ref('a', p) for p in sorted(points))
  tile = ', '.join(['32'] * (dim - 1) + ['*'])
  text = ('kernel: cr%d\nburst width: 64\nunroll factor: 2\niterate: %d\n'
          'input %s: a(%s)\noutput %s: %s = (%s) %s\n' %
          (seed, int(rng.integers(1, 3)), t, tile, t, _ref('b', (0,) * dim),
           refs, '* 0.1f' if t == 'float' else '/ 3'))
  if dim == 2:
    extent = (int(rng.integers(40, 120)), int(rng.integers(20, 40)))
  else:
    extent = (int(rng.integers(40, 90)), int(rng.integers(20, 30)),
              int(rng.integers(16, 24)))
  return text, extent, len(points)
